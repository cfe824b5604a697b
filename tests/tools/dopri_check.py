import sys, importlib, torch
sys.path.insert(0, '.')
import ananke_abm_b200 as ab
from oracle import models_oracle as mo, torchdiffeq_oracle as tdq
oi = importlib.import_module("ananke_abm_b200.odeint")
dev = torch.device('cuda:0')
torch.manual_seed(0)
oracle = mo.OracleModeSep(8); model = ab.ModeSepModel(8, ab.ModeSepConfig()); model.load_state_dict(oracle.state_dict()); model = model.to(dev)
B, T = 200, 7
g = torch.Generator().manual_seed(1)
home = torch.randint(0, 8, (B,), generator=g); work = torch.randint(0, 8, (B,), generator=g); traits = torch.rand(B, 2, generator=g)
t = torch.linspace(0.0, 6.0, T); wgt = torch.linspace(0.5, 1.5, T)[:, None, None]
def rms(a, b): return float((a.double() - b.double()).pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt())
res = {}
def run_oracle(method, tt, sel=None, **kw):
    oracle.zero_grad()
    y0r = oracle.initial_state(home, work, traits).detach().requires_grad_(True)
    ref = tdq.odeint(oracle.rhs, y0r, tt, method=method, **kw)
    if sel is not None: ref = ref[sel]
    ((ref[:, :, :128] * wgt) ** 2).mean().backward()
    return ref.detach(), y0r.grad.clone(), torch.cat([p.grad.reshape(-1) for p in oracle.odefunc.func.net.parameters()])
def run_ours(method, tt, prec, sel=None, fp16=None, **kw):
    for p in model.parameters(): p.grad = None
    y0 = model.initial_state(home.to(dev), work.to(dev), traits.to(dev)).detach().requires_grad_(True)
    opt = {"precision": prec}
    if fp16 is not None: opt["fp16_forward"] = fp16
    out = ab.odeint(model.odefunc, y0, tt.to(dev), method=method, options=opt, **kw)
    if sel is not None: out = out[sel]
    ((out[:, :, :128] * wgt.to(dev)) ** 2).mean().backward()
    return out.detach().cpu(), y0.grad.cpu(), torch.cat([p.grad.reshape(-1) for p in model.odefunc.func.net.parameters()]).cpu()
tf = torch.linspace(0.0, 6.0, 8 * (T - 1) + 1); sel = slice(0, None, 8)
res['oracle rk4 fine'] = run_oracle('rk4', tf, sel)
for tol in (1e-3, 1e-5):
    res[f'oracle dopri5 {tol}'] = run_oracle('dopri5', t, rtol=tol, atol=tol)
res['ours rk4 fine f32'] = run_ours('rk4', tf, 'f32', sel)
res['ours rk4 fine bf16'] = run_ours('rk4', tf, 'bf16', sel)
for fp16 in (False, True):
    for tol in (1e-3, 1e-4, 1e-5, 1e-6):
        res[f'ours dopri5 {"fp16" if fp16 else "bf16"} {tol}'] = run_ours('dopri5', t, 'bf16', fp16=fp16, rtol=tol, atol=tol)
        st = oi._LAST['solver']
        print(f"fwd {'fp16' if fp16 else 'bf16'} tol {tol}: accepted {st.n_accepted} rejected {st.n_rejected} evals {st.n_evals}")
base = res['oracle rk4 fine']
for k, v in res.items():
    print(f"{k:26s} traj {rms(v[0], base[0]):.3e}  gy0 {rms(v[1], base[1]):.3e}  gw {rms(v[2], base[2]):.3e}")
