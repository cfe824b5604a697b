"""What bounds the fused backward stage launch?  Times rk4 forward and rk4 backward (8 steps, CUDA events) with the
wgrad flush optionally disabled; run once per AB200_STAGE_FLAGS value with AB200_STAGE_TIMING_ONLY=1 (16 = no blob spill,
32 = no gx stores: timing experiments, results invalid).   python scripts/bwd_bound.py [B] [nowgrad]"""
import sys, torch
sys.path.insert(0, '.')
import ananke_abm_b200 as ab
from ananke_abm_b200 import stage
dev = torch.device('cuda:0')
torch.manual_seed(42)
m = ab.ModeSepModel(500, ab.ModeSepConfig()).to(dev)
spec = ab.describe_drift(m.odefunc)
w = spec.flat_params().detach()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 189440
nowgrad = len(sys.argv) > 2 and sys.argv[2] == "nowgrad"
y0 = torch.randn(B, 160, device=dev) * 0.3
eng = stage.TcEngine(spec, w)
if nowgrad:
    def _skip(self=eng):
        self.used = 0
    eng.flush = _skip
th = [i * 0.25 for i in range(9)]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
best = [1e9, 1e9]
for it in range(4):
    ev[0].record()
    yp, saved = stage.rk4_forward(eng, y0, th, True)
    ev[1].record()
    gy0, gw = stage.rk4_backward(eng, th, saved, torch.ones_like(yp) / yp.numel())
    ev[2].record()
    torch.cuda.synchronize()
    if it:
        best[0] = min(best[0], ev[0].elapsed_time(ev[1]))
        best[1] = min(best[1], ev[1].elapsed_time(ev[2]))
import os
print(f"flags={os.environ.get('AB200_STAGE_FLAGS', '0')} nowgrad={nowgrad} B={B}: fwd {best[0]:.3f} ms  bwd {best[1]:.3f} ms  (8 steps)")
