"""Per-kernel device times of the tensor-core stage path (CUDA events on the launching stream)."""
import sys, time, torch
sys.path.insert(0, '.')
import ananke_abm_b200 as ab
from ananke_abm_b200 import stage

dev = torch.device('cuda:0')
torch.manual_seed(42)
m = ab.ModeSepModel(500, ab.ModeSepConfig()).to(dev)
spec = ab.describe_drift(m.odefunc)
w = spec.flat_params().detach()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
y0_rm = torch.randn(B, 160, device=dev) * 0.3
y0 = stage.rows_block(y0_rm)
A = [stage.rows_block(torch.randn(B, 64, device=dev) * 0.1) for _ in range(3)]
aout = stage.blocked_zeros(B, 64, dev)
yout = stage.blocked_zeros(B, 160, dev)
eng = stage.TcEngine(spec, w)
dt = 0.25


def timeit(name, fn, reps=10, flops=0.0, bytes_=0.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    h1 = time.perf_counter()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:34s} {ms * 1e3:9.1f} us/launch  host-issue {1e6 * (h1 - h0) / reps:7.1f} us"
          + (f"  {flops / ms / 1e9:8.1f} TFLOP/s" if flops else "") + (f"  {bytes_ / ms / 1e6:8.1f} GB/s" if bytes_ else ""), flush=True)
    return ms


LAYER_FLOP = 2 * (160 * 128 + 4 * 128 * 128 + 128 * 64)   # 188,416 useful flop per agent-eval (no K padding)
timeit("stage_fwd n_a=0 a_out", lambda: eng.stage_forward(y0, [], stage.RK38.stage_input(0, dt), 1.0, B, a_out=aout), flops=B * LAYER_FLOP)
timeit("stage_fwd n_a=3 y_out", lambda: eng.stage_forward(y0, A, stage.RK38.stage_input(3, dt), 1.0, B, y_out=yout,
                                                           cout=stage.RK38.combo(stage.RK38.b, dt)), flops=B * LAYER_FLOP)
eng.backward_begin(B, 4)
G_y0 = stage.blocked_zeros(B, 160, dev)
G_a = [stage.blocked_zeros(B, 64, dev) for _ in range(4)]
lam = stage.rows_block(torch.randn(B, 160, device=dev))
GX = [stage.blocked_zeros(B, 160, dev) for _ in range(4)]


def bwd_stage(i):
    eng.used = 0
    eng.stage_backward(y0, A[:i], stage.RK38.stage_input(i, dt), 1.0, B, G_a[i], GX[i + 1:], [0.1] * (3 - i), [0.2] * (3 - i), GX[i])


BWD_FLOP = 2 * (176 * 128 + 4 * 144 * 128) + 2 * (64 * 128 + 4 * 128 * 128 + 128 * 160)
timeit("stage_bwd n_a=3", lambda: bwd_stage(3), flops=B * BWD_FLOP, bytes_=B * 3040.0)
timeit("stage_bwd n_a=0", lambda: bwd_stage(0), flops=B * BWD_FLOP, bytes_=B * 3040.0)
timeit("combine_backward (init)", lambda: eng.combine_backward(lam, stage.RK38.combo(stage.RK38.b, dt), B, G_y0, G_a, False),
       bytes_=B * (640 + 640 + 4 * 256.0))


def wgrad1():
    eng.used = eng.ntiles
    eng.flush()


def wgrad4():
    eng.used = 4 * eng.ntiles
    eng.flush()


timeit("wgrad 1 stage of blobs", wgrad1, flops=B * LAYER_FLOP, bytes_=B * 3040.0)
timeit("wgrad 4 stages of blobs", wgrad4, flops=4 * B * LAYER_FLOP, bytes_=4 * B * 3040.0)
eng.check_status()

# monolithic forward for comparison
T = 9
t = torch.linspace(0, 2, T, device=dev)
with torch.no_grad():
    timeit("monolithic rk4_tc fwd (8 steps)", lambda: ab.odeint(m.odefunc, y0_rm, t, method='rk4', options={'precision': 'bf16'}),
           reps=3, flops=B * 8 * 4 * LAYER_FLOP)
th = [float(v) for v in t.tolist()]
timeit("rows_block [B,160]", lambda: stage.rows_block(y0_rm, y0), bytes_=B * 1280.0)
timeit("rows_unblock [B,160]", lambda: stage.rows_unblock(y0, B, 160, out=y0_rm), bytes_=B * 1280.0)
timeit("stage rk4 fwd (8 steps, saves a)", lambda: stage.rk4_forward(eng, y0_rm, th, True), reps=3, flops=B * 8 * 4 * LAYER_FLOP)
yp, saved = stage.rk4_forward(eng, y0_rm, th, True)
g = torch.ones_like(yp)
timeit("stage rk4 bwd (8 steps)", lambda: stage.rk4_backward(eng, th, saved, g), reps=3, flops=B * 8 * 4 * (BWD_FLOP + LAYER_FLOP))
print("done")
