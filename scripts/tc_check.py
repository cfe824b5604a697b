import sys, time, torch
sys.path.insert(0, '.')
import ananke_abm_b200 as ab
dev = torch.device('cuda:0')
torch.manual_seed(42)
m = ab.ModeSepModel(500, ab.ModeSepConfig()).to(dev)
def run(B, T, prec):
    g = torch.Generator().manual_seed(1)
    home = torch.randint(0, 500, (B,), generator=g).to(dev); work = torch.randint(0, 500, (B,), generator=g).to(dev)
    traits = torch.rand(B, 2, generator=g).to(dev)
    t = torch.linspace(0, 24, T, device=dev)
    with torch.no_grad():
        y0 = m.initial_state(home, work, traits)
        out = ab.odeint(m.odefunc, y0, t, method='rk4', options={'precision': prec})
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = ab.odeint(m.odefunc, y0, t, method='rk4', options={'precision': prec})
        e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / 3
for (B, T) in [(200, 5), (10000, 97), (148 * 128 * 4, 97)]:
    of, tf = run(B, T, 'f32') if B <= 20000 else (None, None)
    ob, tb = run(B, T, 'bf16')
    asteps = B * (T - 1)
    line = f"B={B} T={T} bf16 {tb:.3f} ms  {asteps / tb * 1e3:.3e} agent-steps/s  {asteps * 755712 / tb * 1e3 / 1e12:.1f} TFLOP/s"
    if of is not None:
        d = (ob - of).abs()
        scale = of.abs().max()
        line += f" | f32 {tf:.3f} ms | max abs err {float(d.max()):.3e} (scale {float(scale):.3g}) rel {float(d.max() / scale):.3e} rms-rel {float(d.pow(2).mean().sqrt() / of.pow(2).mean().sqrt()):.3e}"
        line += f" | row0 equal {bool(torch.equal(ob[0], of[0]))} nan {bool(torch.isnan(ob).any())}"
        pf, lf, _ = m.head(of[:, :2000]); pb, lb, _ = m.head(ob[:, :2000])
        line += f" | label agreement {float((lf.argmax(-1) == lb.argmax(-1)).float().mean()):.4f}"
    print(line, flush=True)
