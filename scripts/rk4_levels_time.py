"""rk4 training step on the tensor-core path (8 grid steps, fwd + bwd) at the three saved_operands levels; CUDA events.
usage: python scripts/rk4_levels_time.py [B]"""
import sys, torch
sys.path.insert(0, '.')
import bench
import ananke_abm_b200 as ab
B = int(sys.argv[1]) if len(sys.argv) > 1 else 250112
dev = torch.device('cuda:0')
cfg = dict(bench.WORKLOADS["c3"], B=B, method="rk4")
model, zfeat, csr = bench.build_model(cfg, "bf16", dev)
home, work, traits, _ = (x.to(dev) for x in bench.make_inputs(cfg, seed=42))
t = torch.linspace(0.0, 2.0, 9, device=dev)
with torch.no_grad():
    table, zemb = model.zone_tables(zfeat, csr)
    y0b = model.initial_state(table, zemb, home, work, traits).contiguous()
for level in ("none", "inputs", "all"):
    ts = []
    for it in range(4):
        y0 = y0b.clone().requires_grad_(True)
        model.zero_grad(set_to_none=True)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        e[0].record()
        out = ab.odeint(model.odefunc, y0, t, method="rk4", options={"precision": "bf16", "saved_operands": level})
        e[1].record()
        out.backward(out.detach() * 1e-3)
        e[2].record()
        torch.cuda.synchronize()
        ts.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
        del out, y0
    f, b = min(x[0] for x in ts[1:]), min(x[1] for x in ts[1:])
    print(f"{level:7s}: forward {f / 8:7.3f} ms per step, backward {b / 8:7.3f} ms per step, sum {(f + b) / 8:7.3f}  (B = {B}, peak {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB)")
