"""One configs[2] training step (GAT zone tables -> initial state -> dopri5 rtol=atol=1e-5 forward -> trajectory loss -> discrete
adjoint) over ONE chunk of agents, for ncu launch lists / captures.  usage: python scripts/prof_c3_step.py [B] [reps] [all|inputs|none]"""
import sys, time, torch
sys.path.insert(0, '.')
import bench
import ananke_abm_b200 as ab
import importlib
oi = importlib.import_module("ananke_abm_b200.odeint")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 333440
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
saved = sys.argv[3] if len(sys.argv) > 3 else "all"
kineto = len(sys.argv) > 4 and sys.argv[4] == "kineto"      # per-kernel durations of the LAST repetition through CUPTI (warm caches, unlike ncu)
dev = torch.device('cuda:0')
cfg = dict(bench.WORKLOADS["c3"], B=B)
model, zfeat, csr = bench.build_model(cfg, "bf16", dev)
model.config.saved_operands = saved
home, work, traits, t = (x.to(dev) for x in bench.make_inputs(cfg, seed=42))
params = list(model.parameters())
prof = None
for it in range(reps):
    if kineto and it == reps - 1:
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
    st = oi._LAST["solver"]
    print(f"rep {it} [{saved}]: {1e3 * (time.perf_counter() - t0):.1f} ms, peak {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, accepted {st.n_accepted} rejected {st.n_rejected}, loss {float(loss):.6g}")
    del y_path, loss
if prof is not None:
    prof.__exit__(None, None, None)
    rows = sorted(((e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0), key=lambda r: -r[2])
    tot = sum(r[2] for r in rows)
    print(f"kernel time of the last repetition: {tot / 1e3:.1f} ms")
    for k, n, us in rows[:16]:
        print(f"  {100 * us / tot:5.1f} %  {us / 1e3:8.2f} ms  {n:4d} x {us / n:8.1f} us  {k[:90]}")
    # GPU idle time: gaps between consecutive kernels on the device timeline, attributed to the kernel that ran BEFORE the gap
    from torch.autograd import DeviceType
    ev = sorted((e for e in prof.events() if e.device_type == DeviceType.CUDA and e.time_range.end > e.time_range.start),
                key=lambda e: e.time_range.start)
    span = ev[-1].time_range.end - ev[0].time_range.start
    gaps = {}
    end = ev[0].time_range.end
    prev = ev[0].name
    for e in ev[1:]:
        g = e.time_range.start - end
        if g > 0:
            gaps.setdefault(prev[:60], [0, 0.0])
            gaps[prev[:60]][0] += 1
            gaps[prev[:60]][1] += g
        if e.time_range.end > end:
            end, prev = e.time_range.end, e.name
    tot_gap = sum(v[1] for v in gaps.values())
    print(f"device timeline: span {span / 1e3:.1f} ms, idle {tot_gap / 1e3:.2f} ms")
    for k, (n, us) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:10]:
        print(f"   idle after {k:60s} {n:4d} gaps {us / 1e3:7.2f} ms  ({us / n:6.1f} us each)")
print("ok")
