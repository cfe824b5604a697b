"""Cycle trace of the forward stage kernel (library built with -DAB200_STAGE_TRACE)."""
import sys, ctypes as C, torch
sys.path.insert(0, '.')
import ananke_abm_b200 as ab
from ananke_abm_b200 import stage
dev = torch.device('cuda:0')
torch.manual_seed(42)
m = ab.ModeSepModel(500, ab.ModeSepConfig()).to(dev)
spec = ab.describe_drift(m.odefunc); w = spec.flat_params().detach()
B = 148 * 128 * 2 * 4
y0 = stage.rows_block(torch.randn(B, 160, device=dev) * 0.3)
aout = stage.blocked_zeros(B, 64, dev)
eng = stage.TcEngine(spec, w)
L = ab.lib()
buf = (C.c_longlong * 8192)(); cnt = (C.c_int * 2)()
for it in range(3):
    eng.stage_forward(y0, [], stage.RK38.stage_input(0, 0.25), 1.0, B, a_out=aout)
    torch.cuda.synchronize()
    L.ab200_debug_stage_trace(buf, cnt)
names = {31: 'i:enter', 32: 'i:fenced', 33: 'i:ready', 34: 'i:mma0', 35: 'i:mmaN', 36: 'i:commit', 1: 'enter', 2: 'st_wait+fence', 3: 'slot_sync', 4: 'issued', 5: 'mma_done', 9: 'tile_start', 10: 'prologue_done', 11: 'net_done'}
for slot in range(2):
    n = min(cnt[slot], 2048)
    ev = [(buf[(slot * 2048 + i) * 2], buf[(slot * 2048 + i) * 2 + 1]) for i in range(n)]
    t0 = ev[0][1]
    print(f"slot {slot}: {n} events")
    prev = t0
    for tag, t in ev[40:110]:
        print(f"   {names.get(tag, tag):14s} t={t - t0:8d}  +{t - prev:6d}")
        prev = t
