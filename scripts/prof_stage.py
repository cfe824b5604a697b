"""Short launch sequence of the tensor-core stage kernels for ncu (one of each kernel, B from argv)."""
import sys, torch
sys.path.insert(0, '.')
import ananke_abm_b200 as ab
from ananke_abm_b200 import stage
dev = torch.device('cuda:0')
torch.manual_seed(42)
m = ab.ModeSepModel(500, ab.ModeSepConfig()).to(dev)
spec = ab.describe_drift(m.odefunc)
w = spec.flat_params().detach()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 4
y0 = stage.rows_block(torch.randn(B, 160, device=dev) * 0.3)
A = [stage.rows_block(torch.randn(B, 64, device=dev) * 0.1) for _ in range(3)]
aout = stage.blocked_zeros(B, 64, dev); yout = stage.blocked_zeros(B, 160, dev)
eng = stage.TcEngine(spec, w)
dt = 0.25
eng.backward_begin(B, 4)
G_y0 = stage.blocked_zeros(B, 160, dev); G_a = [stage.blocked_zeros(B, 64, dev) for _ in range(4)]
GX = [stage.blocked_zeros(B, 160, dev) for _ in range(4)]
for it in range(3):
    eng.stage_forward(y0, [], stage.RK38.stage_input(0, dt), 1.0, B, a_out=aout)
    eng.stage_forward(y0, A, stage.RK38.stage_input(3, dt), 1.0, B, y_out=yout, cout=stage.RK38.combo(stage.RK38.b, dt))
    eng.used = 0
    for i in (3, 2, 1, 0):
        eng.stage_backward(y0, A[:i], stage.RK38.stage_input(i, dt), 1.0, B, G_a[i], GX[i + 1:], [0.1] * (3 - i), [0.2] * (3 - i), GX[i])
    eng.flush()
torch.cuda.synchronize()
eng.check_status()
print("ok")
