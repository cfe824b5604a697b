"""CUDA-event timing of the fused embedding-space loss pass (forward, forward + backward) on a [B, T, 64] slab."""
import sys, torch
sys.path.insert(0, '.')
import ananke_abm_b200 as ab
dev = torch.device('cuda:0')
B, T, Z = (int(sys.argv[1]) if len(sys.argv) > 1 else 189_440), 16, 10_000
torch.manual_seed(0)
emb = (0.3 * torch.randn(B, T, 64, device=dev)).requires_grad_(True)
vel = (0.3 * torch.randn(B, T, 64, device=dev)).requires_grad_(True)
tab = (0.3 * torch.randn(Z, 64, device=dev)).requires_grad_(True)
r = torch.rand(B, T, device=dev)
is_gt, stay, trav = r < 0.12, (r >= 0.12) & (r < 0.6), r >= 0.6
zid = torch.randint(0, Z, (B, T), device=dev)
neg = torch.full_like(zid, -1)
y_gt, y_st, prev, dest = torch.where(is_gt, zid, neg), torch.where(stay, zid, neg), torch.where(trav, zid, neg), torch.where(trav, (zid * 7 + 3) % Z, neg)
def fwd():
    return ab.emb_loss_terms(emb, vel, tab, y_gt, is_gt, y_st, stay, trav, prev, dest, is_gt)
def both():
    for x in (emb, vel, tab):
        x.grad = None
    sum(fwd().values()).backward()
for name, fn in (("forward", fwd), ("forward+backward", both)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    rows = B * T
    print(f"{name}: {ms * 1e3:.0f} us for {rows} rows ({rows * 512 / ms / 1e6:.0f} GB/s of pred_emb + v reads alone)")
