import sys, numpy as np, torch
sys.path.insert(0, '.')
import ananke_abm_b200 as ab
from oracle import models_oracle as mo, torchdiffeq_oracle as tdq
from ananke_abm_b200.drift import describe_drift
dev = torch.device('cuda:0')
g = np.load('tests/golden/rhs_fixture.npz')
m = ab.ModeSepModel(8, ab.ModeSepConfig())
m.load_state_dict({k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('ms_sd_')})
m = m.to(dev)
print('spec', describe_drift(m.odefunc) is not None)
y = torch.from_numpy(g['ms_y']).to(dev)
f = m.odefunc(torch.tensor(float(g['ms_t0'])), y).cpu()
ref = torch.from_numpy(g['ms_f0'])
d = (f - ref).abs()
print('max err', float(d.max()), 'ref max', float(ref.abs().max()))
print('err by col block (p,v,h):', float(d[:, :64].max()), float(d[:, 64:128].max()), float(d[:, 128:].max()))
print('err per agent:', [round(float(x), 4) for x in d[:, 64:128].max(dim=1).values])
print('err per col:', [round(float(x), 4) for x in d[:, 64:128].max(dim=0).values])
# bigger batch via oracle
torch.manual_seed(0)
orc = mo.OracleModeSep(8); orc.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
yy = torch.randn(200, 160) * 0.4
fr = orc.rhs(torch.tensor(3.0), yy).detach()
fg = m.odefunc(torch.tensor(3.0), yy.to(dev)).cpu()
d = (fg - fr).abs()
print('B=200 err', float(d.max()), 'per-agent-block of 64:', [round(float(d[i:i+64].max()), 5) for i in range(0, 200, 64)])
# rk4 path
t = torch.linspace(0, 5, 6)
y0 = yy[:70].clone().requires_grad_(True)
refp = tdq.odeint(orc.rhs, y0, t, method='rk4')
(refp[:, :, :128] ** 2).mean().backward()
y0d = yy[:70].to(dev).clone().requires_grad_(True)
out = ab.odeint(m.odefunc, y0d, t.to(dev), method='rk4')
print('out.requires_grad', out.requires_grad, out.grad_fn)
(out[:, :, :128] ** 2).mean().backward()
print('rk4 fwd err', float((out.detach().cpu() - refp.detach()).abs().max()), 'ref max', float(refp.abs().max()))
print('gy0 err', float((y0d.grad.cpu() - y0.grad).abs().max()), 'ref max', float(y0.grad.abs().max()))
for (n, p), (_, q) in zip(m.odefunc.func.net.named_parameters(), orc.odefunc.func.net.named_parameters()):
    print(n, None if p.grad is None else float((p.grad.cpu() - q.grad).abs().max()), float(q.grad.abs().max()))
