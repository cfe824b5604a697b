import sys, ctypes as C, torch
sys.path.insert(0, '.')
import ananke_abm_b200 as ab
L = ab.lib()
dev = torch.device('cuda:0')
torch.manual_seed(0)
for (N, K) in [(128, 128), (64, 128), (128, 64), (128, 192)]:
    A = torch.randn(128, K, device=dev)
    B = torch.randn(N, K, device=dev)
    ref = (A.bfloat16().float() @ B.bfloat16().float().T)
    for a_mode in (0, 1, 2, 3, 4):
        for b_mode in (1, 2, 3, 4):
            D = torch.zeros(128, N, device=dev)
            st = torch.full((1,), -7, dtype=torch.int32, device=dev)
            rc = L.ab200_debug_umma_probe(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, a_mode, b_mode, st.data_ptr(), torch.cuda.current_stream().cuda_stream)
            try:
                torch.cuda.synchronize()
                err = float((D - ref).abs().max())
                print(f"N={N} K={K} a_mode={a_mode} b_mode={b_mode} rc={rc} status={int(st)} maxerr={err:.4g} refmax={float(ref.abs().max()):.3g}", flush=True)
            except Exception as e:
                print(f"N={N} K={K} a_mode={a_mode} b_mode={b_mode} rc={rc} EXC {e}", flush=True)
                sys.exit(1)
