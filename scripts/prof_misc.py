"""Every non-tensor-core kernel of the path once, at the configs[2] shapes, for one `ncu --set full` capture:
GAT forward/backward (Z = 10k, 4 heads), fused cosine head + argmax, generic-func stage combine / error norm, and the
blocked-layout elementwise kernels of one dopri5 forward+backward chunk.   python scripts/prof_misc.py [agents]"""
import sys, torch
sys.path.insert(0, '.')
import ananke_abm_b200 as ab
from ananke_abm_b200.graph import synthetic_zone_graph
from ananke_abm_b200.inference import head_argmax
import importlib
oi = importlib.import_module("ananke_abm_b200.odeint")

dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 189_440
Z = 10_000
torch.manual_seed(42)
mc = ab.ModeSepConfig()
mc.precision = "bf16"
mc.ode_method = "dopri5"
model = ab.GATODEModel(7, mc, heads=4).to(dev)
ei, feats = synthetic_zone_graph(Z, k=6, seed=42)
csr = ab.build_zone_csr(ei, Z).to(dev)
feats = feats.to(dev)
g = torch.Generator().manual_seed(1)
home = torch.randint(0, Z, (B,), generator=g).to(dev)
work = torch.randint(0, Z, (B,), generator=g).to(dev)
traits = torch.rand(B, 2, generator=g).to(dev)

for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    # 1. GAT forward + backward (both zone tables)
    for p in model.parameters():
        p.grad = None
    table, zemb = model.zone_tables(feats, csr)
    (table.square().mean() + zemb.square().mean()).backward()
    # 2. fused head: nomination on tensor cores + fp32 re-score (configs[1]/[2] label prediction)
    with torch.no_grad():
        table, zemb = model.zone_tables(feats, csr)
        emb = torch.randn(min(B, 65_536), 8, 64, device=dev)
        labels = head_argmax(emb, table, mc.softmax_tau)
    # 3. generic func: fused stage combine + error norm (fp32 elementwise, HBM-bound)
    with torch.no_grad():
        yg = torch.randn(B, 160, device=dev)
        oi.odeint(lambda t, y: -0.1 * y, yg, torch.tensor([0.0, 0.5], device=dev), method="dopri5", rtol=1e-5, atol=1e-5)
    # 4. one dopri5 chunk forward + backward on the stage path (elementwise kernels around the tensor-core launches)
    table, zemb = model.zone_tables(feats, csr)
    y0 = model.initial_state(table, zemb, home, work, traits)
    yp = model.integrate(y0, torch.tensor([0.0, 0.25, 0.5], device=dev))
    yp[:, :, :128].square().mean().backward()
    # 5. the fused embedding-space loss terms (forward + backward) on a [B, 16, 64] slab with random stay / travel structure
    T2 = 16
    emb = (0.3 * torch.randn(B, T2, 64, device=dev)).requires_grad_(True)
    vel = (0.3 * torch.randn(B, T2, 64, device=dev)).requires_grad_(True)
    tab = table.detach().clone().requires_grad_(True)
    r = torch.rand(B, T2, device=dev)
    is_gt, stay, trav = r < 0.12, (r >= 0.12) & (r < 0.6), r >= 0.6
    zid = torch.randint(0, Z, (B, T2), device=dev)
    y_gt = torch.where(is_gt, zid, torch.full_like(zid, -1))
    y_st = torch.where(stay, zid, torch.full_like(zid, -1))
    prev = torch.where(trav, zid, torch.full_like(zid, -1))
    dest = torch.where(trav, (zid * 7 + 3) % Z, torch.full_like(zid, -1))
    terms = ab.emb_loss_terms(emb, vel, tab, y_gt, is_gt, y_st, stay, trav, prev, dest, is_gt)
    sum(terms.values()).backward()
    # 6. strict-fp32 drift evaluation and its vector-Jacobian product (FFMA kernels behind autograd-through-dopri5 / the adjoint)
    spec = ab.describe_drift(model.odefunc)
    yv = torch.randn(min(B, 65_536), 160, device=dev)
    fv = oi.drift_eval(spec, spec.flat_params().detach(), 3.0, yv)
    gyv, gwv = oi.drift_vjp(spec, spec.flat_params().detach(), 3.0, yv, fv)
    torch.cuda.synchronize()
print("ok", int(labels.sum()) >= 0)
