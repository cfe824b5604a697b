"""GPU check of the tensor-core stage path against the strict-fp32 kernels (run under gpurun)."""
import sys, time, torch
sys.path.insert(0, '.')
import ananke_abm_b200 as ab
from ananke_abm_b200 import stage, _lib
from ananke_abm_b200.odeint import drift_eval

dev = torch.device('cuda:0')
torch.manual_seed(42)
m = ab.ModeSepModel(500, ab.ModeSepConfig()).to(dev)
spec = ab.describe_drift(m.odefunc)
w = spec.flat_params().detach()


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def rms_rel(a, b):
    return float((a.double() - b.double()).pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt().clamp_min(1e-30))


def inputs(B):
    g = torch.Generator().manual_seed(1)
    home = torch.randint(0, 500, (B,), generator=g).to(dev); work = torch.randint(0, 500, (B,), generator=g).to(dev)
    traits = torch.rand(B, 2, generator=g).to(dev)
    with torch.no_grad():
        y0 = m.initial_state(home, work, traits)
        y0[:, 64:128] = 0.1 * torch.randn(B, 64, device=dev)
    return y0.contiguous()


# 1. single stage
for B in (100, 1000, 4096 + 17):
    y0 = inputs(B)
    eng = stage.TcEngine(spec, w)
    ab_ = stage.blocked_zeros(B, 64, dev)
    eng.stage_forward(stage.rows_block(y0), [], stage.Combo(0.0, [], []), 3.7, B, a_out=ab_)
    a_out = stage.rows_unblock(ab_, B, 64)
    torch.cuda.synchronize()
    eng.check_status()
    ref = drift_eval(spec, w, 3.7, y0)[:, 64:128]
    print(f"stage fwd B={B}: rel {rel(a_out, ref):.3e} rms {rms_rel(a_out, ref):.3e} nan {bool(torch.isnan(a_out).any())}", flush=True)

# 2./3. rk4 forward + backward vs the fp32 kernels
for (B, T) in ((300, 5), (2000, 9)):
    y0 = inputs(B)
    t = torch.linspace(0.0, 2.0, T, device=dev)
    outs = {}
    for prec in ('f32', 'bf16'):
        for p in m.odefunc.parameters():
            p.grad = None
        y = y0.clone().requires_grad_(True)
        yp = ab.odeint(m.odefunc, y, t, method='rk4', options={'precision': prec})
        wgt = torch.linspace(0.5, 1.5, T, device=dev)[:, None, None]
        ((yp[:, :, :128] * wgt) ** 2).mean().backward()
        torch.cuda.synchronize()
        outs[prec] = (yp.detach(), y.grad.clone(), torch.cat([p.grad.reshape(-1) for p in m.odefunc.parameters()]))
    f, b = outs['f32'], outs['bf16']
    print(f"rk4 B={B} T={T}: y_path rel {rel(b[0], f[0]):.3e} rms {rms_rel(b[0], f[0]):.3e} | gy0 rel {rel(b[1], f[1]):.3e} rms {rms_rel(b[1], f[1]):.3e}"
          f" | gw rel {rel(b[2], f[2]):.3e} rms {rms_rel(b[2], f[2]):.3e} nan {bool(torch.isnan(b[2]).any())}", flush=True)
    # per-parameter-block error
    off = 0
    for name, p in m.odefunc.named_parameters():
        n = p.numel()
        print(f"    {name:28s} rms-rel {rms_rel(b[2][off:off + n], f[2][off:off + n]):.3e}  rel {rel(b[2][off:off + n], f[2][off:off + n]):.3e}")
        off += n

# 4. timing at scale
B, T = 148 * 128 * 4, 25
y0 = inputs(B)
t = torch.linspace(0.0, 6.0, T, device=dev)
th = [float(v) for v in t.tolist()]
eng = stage.TcEngine(spec, w)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    yp, saved = stage.rk4_forward(eng, y0, th, save_stages=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    g = torch.ones_like(yp) / yp.numel()
    gy0, gw = stage.rk4_backward(eng, th, saved, g)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    eng.check_status()
    n = B * (T - 1)
    print(f"B={B} T={T}: stage fwd {1e3 * (t1 - t0):.1f} ms ({n / (t1 - t0):.3e} agent-steps/s, {n * 755712 / (t1 - t0) / 1e12:.0f} TFLOP/s)"
          f" | bwd {1e3 * (t2 - t1):.1f} ms ({n / (t2 - t1):.3e} agent-steps/s) | fwd+bwd {n / (t2 - t0):.3e}", flush=True)
print("done")
