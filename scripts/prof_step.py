"""One rk4 training step sequence on the stage path for ncu (B from argv)."""
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
y0 = torch.randn(B, 160, device=dev) * 0.3
eng = stage.TcEngine(spec, w)
th = [0.0, 0.25, 0.5, 0.75]
for it in range(2):
    yp, saved = stage.rk4_forward(eng, y0, th, True)
    gy0, gw = stage.rk4_backward(eng, th, saved, torch.ones_like(yp) / yp.numel())
torch.cuda.synchronize()
eng.check_status()
print("ok")
