"""Fused cross-entropy head timing (CUDA events): forward (log-sum-exp + target logit) and backward (two tensor-core
passes) at M rows x Z zones.   python scripts/head_ce_time.py [M] [Z]"""
import sys, torch
sys.path.insert(0, '.')
import ananke_abm_b200 as ab
dev = torch.device('cuda:0')
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
Z = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
g = torch.Generator().manual_seed(0)
emb = torch.randn(M, 64, generator=g).to(dev).requires_grad_(True)
table = torch.randn(Z, 64, generator=g).to(dev).requires_grad_(True)
tgt = torch.randint(0, Z, (M,), generator=g).to(dev)
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
bf = bb = 1e9
for it in range(4):
    emb.grad = table.grad = None
    e[0].record()
    loss = ab.head_ce_rows(emb, table, tgt, 0.2).mean()
    e[1].record()
    loss.backward()
    e[2].record()
    torch.cuda.synchronize()
    if it:
        bf, bb = min(bf, e[0].elapsed_time(e[1])), min(bb, e[1].elapsed_time(e[2]))
fl = 2.0 * M * Z * 64
print(f"M={M} Z={Z}: loss {float(loss):.6f}; forward {bf:.2f} ms ({fl / bf / 1e9:.0f} algorithmic TFLOP/s); "
      f"backward {bb:.2f} ms ({2 * fl / bb / 1e9:.0f} algorithmic TFLOP/s of the two gradient contractions)")
# with the expected-distance term (rows in target order inside head_loss_rows)
xy = torch.rand(Z, 2, generator=g)
dist = torch.cdist(xy, xy).to(dev)
bf = bb = 1e9
for it in range(3):
    emb.grad = table.grad = None
    e[0].record()
    ce, ed = ab.head_loss_rows(emb, table, tgt, 0.2, dist)
    loss = ce.mean() + 0.5 * ed.mean()
    e[1].record()
    loss.backward()
    e[2].record()
    torch.cuda.synchronize()
    if it:
        bf, bb = min(bf, e[0].elapsed_time(e[1])), min(bb, e[1].elapsed_time(e[2]))
print(f"  + expected distance: forward {bf:.2f} ms, backward {bb:.2f} ms (including the sort / gather by target)")
