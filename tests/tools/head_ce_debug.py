"""Where does the tensor-core CE backward differ from the float64 reference?  (dev tool; run from the repo root)"""
import sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
import ananke_abm_b200 as ab
dev = torch.device('cuda:0')
M, Z = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator().manual_seed(M * 7 + Z)
emb = torch.randn(M, 64, generator=g).to(dev).requires_grad_(True)
table = torch.randn(Z, 64, generator=g).to(dev).requires_grad_(True)
tgt = torch.randint(0, Z, (M,), generator=g).to(dev)
w = torch.rand(M, generator=g).to(dev)
try:
    (ab.head_ce_rows(emb, table, tgt, 0.2) * w).sum().backward()
except Exception as ex:
    print("EXC", ex); sys.exit(0)
e2, t2 = emb.detach().double().requires_grad_(True), table.detach().double().requires_grad_(True)
tn = t2 / (t2.norm(dim=-1, keepdim=True) + 1e-8); en = e2 / (e2.norm(dim=-1, keepdim=True) + 1e-8)
(F.cross_entropy(en @ tn.T / 0.2, tgt, reduction="none") * w.double()).sum().backward()
for name, a, b in (("emb", emb.grad, e2.grad), ("table", table.grad, t2.grad)):
    err = (a.double() - b).abs().amax(dim=1) / b.abs().max()
    bad = (err > 3e-5).nonzero().flatten()
    print(name, "max rel err", float(err.max()), "bad rows", bad.numel(), "of", err.numel(),
          "tiles:", sorted(set((bad // 128).tolist()))[:40])
