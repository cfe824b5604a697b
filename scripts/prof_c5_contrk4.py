"""One configs[4] step in the continuous-rk4 mode (bench.py --workload c5 --adjoint-mode continuous-rk4): GAT zone tables -> initial
state -> rk4 forward (step_size 0.25 over t = [0, 24], two rows kept) -> loss -> continuous adjoint on the tensor-core stage kernels.
Per-kernel durations of the last repetition through CUPTI.  usage: python scripts/prof_c5_contrk4.py [B] [reps]"""
import sys, time, types, torch
sys.path.insert(0, '.')
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device('cuda:0')
args = types.SimpleNamespace(workload="c5", agents=B, solver="", adjoint_mode="continuous-rk4")
cfg = bench._config_for(args)
model, zfeat, csr = bench.build_model(cfg, "bf16", dev, "all", "continuous-rk4")
if len(sys.argv) > 3:      # launch structure of the backward pass: fused | staged (adjoint_tc.rk4_continuous_adjoint)
    model.config.adjoint_fused = {"fused": True, "staged": False}[sys.argv[3]]
home, work, traits, t = (x.to(dev) for x in bench.make_inputs(cfg, seed=42))
params = list(model.parameters())
prof = None
for it in range(reps):
    if it == reps - 1:
        from torch.profiler import profile, ProfilerActivity
        prof = profile(activities=[ProfilerActivity.CUDA])
        prof.__enter__()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for p in params:
        p.grad = None
    table, zemb = model.zone_tables(zfeat, csr)
    y0 = model.initial_state(table, zemb, home, work, traits)
    y_path = model.integrate(y0, t)
    loss = bench._TrajectoryLoss.apply(y_path)
    loss.backward()
    torch.cuda.synchronize()
    print(f"rep {it}: {1e3 * (time.perf_counter() - t0):.1f} ms, peak {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, rows {tuple(y_path.shape)}, loss {float(loss):.6g}")
    del y_path, loss
prof.__exit__(None, None, None)
rows = sorted(((e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0), key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"kernel time of the last repetition: {tot / 1e3:.1f} ms")
for k, n, us in rows[:14]:
    print(f"  {100 * us / tot:5.1f} %  {us / 1e3:8.2f} ms  {n:4d} x {us / n:8.1f} us  {k[:100]}")
