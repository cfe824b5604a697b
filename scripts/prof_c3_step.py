"""One configs[2] training step (GAT zone tables -> initial state -> dopri5 rtol=atol=1e-5 forward -> trajectory loss -> discrete
adjoint) over ONE chunk of agents, for ncu launch lists / captures.  usage: python scripts/prof_c3_step.py [B] [reps]"""
import sys, time, torch
sys.path.insert(0, '.')
import bench
import ananke_abm_b200 as ab
import importlib
oi = importlib.import_module("ananke_abm_b200.odeint")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 333440
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device('cuda:0')
cfg = dict(bench.WORKLOADS["c3"], B=B)
model, zfeat, csr = bench.build_model(cfg, "bf16", dev)
home, work, traits, t = (x.to(dev) for x in bench.make_inputs(cfg, seed=42))
params = list(model.parameters())
for it in range(reps):
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
    print(f"rep {it}: {1e3 * (time.perf_counter() - t0):.1f} ms, accepted {st.n_accepted} rejected {st.n_rejected}, loss {float(loss):.6g}")
    del y_path, loss
print("ok")
