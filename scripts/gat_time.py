"""GAT layer timing (CUDA events) at the configs[2] graph (Z = 10k, L2-resident) and at a graph large enough to leave L2
(Z = 1M zones, ~8 nnz/row): achieved GB/s against the algorithmic bytes of SURVEY.md §8(d)
    bytes/eval = 4*(Z*F_in + 2*Z*Hh*F' + 2*Z*Hh) + 4*(Z+1) + 4*nnz.
python scripts/gat_time.py [Z ...]"""
import sys, torch
sys.path.insert(0, '.')
import ananke_abm_b200 as ab
from ananke_abm_b200.graph import synthetic_zone_graph
dev = torch.device('cuda:0')
for Z in [int(a) for a in sys.argv[1:]] or [10_000, 1_000_000]:
    ei, feats = synthetic_zone_graph(Z, k=6, seed=42)
    csr = ab.build_zone_csr(ei, Z).to(dev)
    x = feats.to(dev).requires_grad_(True)
    nnz = int(csr.col.numel())
    for heads, fo in ((4, 16), (1, 8)):
        torch.manual_seed(0)
        gat = ab.GATEmbed(7, fo, heads=heads).to(dev)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        bf = bb = 1e9
        for it in range(6):
            x.grad = None
            e[0].record()
            out = gat(x, csr)
            e[1].record()
            g = torch.autograd.grad(out, [x] + list(gat.parameters()), torch.ones_like(out))
            e[2].record()
            torch.cuda.synchronize()
            if it:
                bf, bb = min(bf, e[0].elapsed_time(e[1])), min(bb, e[1].elapsed_time(e[2]))
        alg = 4 * (Z * 7 + 2 * Z * heads * fo + 2 * Z * heads) + 4 * (Z + 1) + 4 * nnz
        print(f"Z={Z} nnz={nnz} heads={heads} F'={fo}: fwd {bf * 1e3:.1f} us = {alg / bf / 1e6:.0f} GB/s of algorithmic bytes "
              f"({alg / 1e6:.1f} MB); bwd {bb * 1e3:.1f} us (incl. ones_like + autograd dispatch)")
