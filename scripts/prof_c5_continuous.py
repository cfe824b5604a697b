"""Kernel shares of one training step through the odeint_adjoint seam with torchdiffeq's CONTINUOUS adjoint (bench --workload c5
--adjoint-mode continuous) over one chunk of agents.  usage: python scripts/prof_c5_continuous.py [B]"""
import sys, time, torch
sys.path.insert(0, '.')
import bench
import importlib
oi = importlib.import_module("ananke_abm_b200.odeint")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
dev = torch.device('cuda:0')
cfg = dict(bench.WORKLOADS["c5"], B=B)
model, zfeat, csr = bench.build_model(cfg, "bf16", dev, "all", "continuous")
home, work, traits, t = (x.to(dev) for x in bench.make_inputs(cfg, seed=42))
params = list(model.parameters())
from torch.profiler import profile, ProfilerActivity
for it in range(2):
    prof = profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) if it == 1 else None
    if prof: prof.__enter__()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for p in params: p.grad = None
    table, zemb = model.zone_tables(zfeat, csr)
    y0 = model.initial_state(table, zemb, home, work, traits)
    y_path = model.integrate(y0, t)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    loss = bench._TrajectoryLoss.apply(y_path)
    loss.backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"rep {it}: forward {1e3*(t1-t0):.1f} ms, backward {1e3*(t2-t1):.1f} ms, solver evals (last solve) {oi._LAST['solver'].n_evals if hasattr(oi._LAST['solver'],'n_evals') else '?'}")
    if prof:
        prof.__exit__(None, None, None)
        rows = sorted(((e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0), key=lambda r: -r[2])
        tot = sum(r[2] for r in rows)
        print(f"kernel time {tot/1e3:.1f} ms")
        for k, n, us in rows[:22]:
            print(f"  {100*us/tot:5.1f} %  {us/1e3:9.2f} ms  {n:6d} x {us/max(n,1):8.1f} us  {k[:100]}")
        cpu = sorted(((e.key, e.count, e.self_cpu_time_total) for e in prof.key_averages()), key=lambda r: -r[2])[:8]
        for k, n, us in cpu:
            print(f"  cpu {us/1e3:9.2f} ms {n:6d} x  {k[:80]}")
