import sys, torch
sys.path.insert(0, '.')
from ananke_abm_b200.inference import head_argmax
dev = torch.device('cuda:0')
for (M, Z) in ((970000, 500), (1000000, 10000), (4000000, 10000)):
    emb = torch.randn(M, 64, device=dev); table = torch.randn(Z, 64, device=dev)
    for _ in range(2): head_argmax(emb, table)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): lab = head_argmax(emb, table)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * M * Z * 64
    # torch reference, chunked
    def ref():
        tn = table / (table.norm(dim=-1, keepdim=True) + 1e-8)
        outs = []
        for s in range(0, M, 65536):
            e = emb[s:s + 65536]
            outs.append(((e / (e.norm(dim=-1, keepdim=True) + 1e-8)) @ tn.T).argmax(-1))
        return torch.cat(outs)
    for _ in range(2): r = ref()
    torch.cuda.synchronize(); e0.record()
    for _ in range(3): r = ref()
    e1.record(); torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / 3
    print(f"M={M} Z={Z}: fused {ms:.2f} ms ({fl / ms / 1e9:.0f} algorithmic TFLOP/s, {3 * fl / ms / 1e9:.0f} issued) | torch fp32 chunked {ms_ref:.2f} ms | "
          f"labels equal {float((lab == r).float().mean()):.6f}", flush=True)
