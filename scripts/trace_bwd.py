"""Cycle trace of the fused backward stage kernel (library built with AB200_STAGE_TRACE=1)."""
import sys, ctypes as C, torch
sys.path.insert(0, '.')
import ananke_abm_b200 as ab
from ananke_abm_b200 import stage
dev = torch.device('cuda:0')
torch.manual_seed(42)
m = ab.ModeSepModel(500, ab.ModeSepConfig()).to(dev)
spec = ab.describe_drift(m.odefunc); w = spec.flat_params().detach()
B = 148 * 128 * 2 * 3
y0 = torch.randn(B, 160, device=dev) * 0.3
eng = stage.TcEngine(spec, w)
th = [0.0, 0.25, 0.5]
L = ab.lib()
buf = (C.c_longlong * 8192)(); cnt = (C.c_int * 2)()
yp, saved = stage.rk4_forward(eng, y0, th, True)
for it in range(2):
    L.ab200_debug_stage_trace_bwd(buf, cnt)
    gy0, gw = stage.rk4_backward(eng, th, saved, torch.ones_like(yp) / yp.numel())
    torch.cuda.synchronize()
L.ab200_debug_stage_trace_bwd(buf, cnt)
names = {1: 'enter', 2: 'st_wait+fence', 3: 'slot_sync', 4: 'issued', 5: 'mma_done', 9: 'stage_start', 10: 'prologue_done', 12: 'fwd_recompute_done',
         13: 'gO_ready', 11: 'dgrad_done'}
slot = 0
n = min(cnt[slot], 2048)
ev = [(buf[(slot * 2048 + i) * 2], buf[(slot * 2048 + i) * 2 + 1]) for i in range(n)]
# summarise per phase over all stage-tiles
import collections
tot = collections.defaultdict(int); cntp = collections.defaultdict(int)
prev_tag, prev_t = None, None
for tag, t in ev:
    if prev_tag is not None and t > prev_t:
        key = f"{names.get(prev_tag, prev_tag)} -> {names.get(tag, tag)}"
        tot[key] += t - prev_t; cntp[key] += 1
    prev_tag, prev_t = tag, t
total = sum(tot.values())
print(f"slot 0: {n} events, total {total} cycles")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"  {k:42s} {v:9d} cycles  {100.0 * v / total:5.1f}%  n={cntp[k]:4d}  avg {v / cntp[k]:8.0f}")
