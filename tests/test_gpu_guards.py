"""Out-of-bounds guards without compute-sanitizer: every output / workspace of the newer kernels is carved out of a larger
sentinel-filled buffer with exactly the size the C ABI asks for; the sentinels on both sides must survive the call."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
SENT = 0x5A
PAD = 4096


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


class Guarded:
    """nbytes of device memory with PAD sentinel bytes before and after (start 256-byte aligned)."""

    def __init__(self, nbytes: int, dev):
        self.n = int(nbytes)
        self.raw = torch.full((PAD + self.n + PAD + 256,), SENT, dtype=torch.uint8, device=dev)
        off = (-self.raw.data_ptr()) % 256
        self.lo = off + PAD - (PAD % 256)
        self.view = self.raw[self.lo:self.lo + self.n]

    def ptr(self):
        return self.view.data_ptr()

    def as_f32(self, *shape):
        return self.view.view(torch.float32).view(*shape)

    def intact(self) -> bool:
        return bool((self.raw[:self.lo] == SENT).all()) and bool((self.raw[self.lo + self.n:] == SENT).all())


@pytest.mark.parametrize("Z,heads,F_out,concat", [(37, 2, 32, True), (129, 4, 16, True), (130, 4, 16, False), (8, 1, 8, True), (1000, 1, 4, True)])
def test_gat_kernels_stay_inside_their_buffers(Z, heads, F_out, concat):
    from ananke_abm_b200 import _lib
    from ananke_abm_b200.graph import build_zone_csr, synthetic_zone_graph
    dev = _cuda()
    L = _lib.lib()
    ei, x = synthetic_zone_graph(Z, k=6, seed=Z) if Z > 8 else (torch.tensor([[0, 1, 2, 3, 4, 5, 6], [1, 2, 3, 4, 5, 6, 7]]), torch.rand(8, 7))
    csr = build_zone_csr(ei, Z).to(dev)
    x = x.to(dev).contiguous()
    HF, F_in = heads * F_out, 7
    g = torch.Generator().manual_seed(1)
    W = (torch.randn(HF, F_in, generator=g) * 0.3).to(dev)
    a_s, a_d = torch.randn(HF, generator=g).to(dev), torch.randn(HF, generator=g).to(dev)
    bias = torch.randn(HF if concat else F_out, generator=g).to(dev)
    n_out = Z * (HF if concat else F_out)
    bufs = {k: Guarded(4 * n, dev) for k, n in (("out", n_out), ("xw", Z * HF), ("as", Z * heads), ("ad", Z * heads),
                                                 ("alpha", csr.nnz * heads))}
    st = torch.cuda.current_stream().cuda_stream
    rc = L.ab200_gat_forward(csr.rowptr.data_ptr(), csr.col.data_ptr(), Z, csr.nnz, x.data_ptr(), F_in, W.data_ptr(), a_s.data_ptr(),
                             a_d.data_ptr(), bias.data_ptr(), heads, F_out, 1 if concat else 0, 0.2, bufs["out"].ptr(), bufs["xw"].ptr(),
                             bufs["as"].ptr(), bufs["ad"].ptr(), bufs["alpha"].ptr(), st)
    _lib.check(rc, "ab200_gat_forward")
    torch.cuda.synchronize()
    assert all(b.intact() for b in bufs.values())
    gout = torch.randn(n_out, generator=g).to(dev)
    ws = Guarded(L.ab200_gat_backward_workspace_bytes(Z, csr.nnz, heads, F_out), dev)
    outs = {k: Guarded(4 * n, dev) for k, n in (("gx", Z * F_in), ("gW", HF * F_in), ("gas", HF), ("gad", HF),
                                                 ("gb", HF if concat else F_out))}
    rc = L.ab200_gat_backward(csr.rowptr.data_ptr(), csr.col.data_ptr(), csr.rowptr_t.data_ptr(), csr.col_t.data_ptr(),
                              csr.eid_t.data_ptr(), Z, csr.nnz, x.data_ptr(), F_in, W.data_ptr(), a_s.data_ptr(), a_d.data_ptr(), heads,
                              F_out, 1 if concat else 0, 0.2, bufs["xw"].ptr(), bufs["as"].ptr(), bufs["ad"].ptr(), bufs["alpha"].ptr(),
                              gout.data_ptr(), outs["gx"].ptr(), outs["gW"].ptr(), outs["gas"].ptr(), outs["gad"].ptr(), outs["gb"].ptr(),
                              ws.ptr(), ws.n, st)
    _lib.check(rc, "ab200_gat_backward")
    torch.cuda.synchronize()
    assert ws.intact() and all(b.intact() for b in outs.values()) and all(b.intact() for b in bufs.values())
    assert all(torch.isfinite(b.view.view(torch.float32)).all() for b in outs.values())


@pytest.mark.parametrize("M,Z,with_dist", [(1, 8, False), (130, 129, True), (300, 20_000, False), (20_000, 300, True), (4097, 1000, True)])
def test_head_loss_kernels_stay_inside_their_buffers(M, Z, with_dist):
    from ananke_abm_b200 import _lib
    dev = _cuda()
    L = _lib.lib()
    g = torch.Generator().manual_seed(M + Z)
    emb, table = torch.randn(M, 64, generator=g).to(dev), torch.randn(Z, 64, generator=g).to(dev)
    tgt = torch.sort(torch.randint(0, Z, (M,), generator=g)).values.to(dev)
    dist = torch.rand(Z, Z, generator=g).to(dev) if with_dist else None
    st = torch.cuda.current_stream().cuda_stream
    ws = Guarded(L.ab200_head_workspace_bytes(Z, 64), dev)
    lse, tl, ed = Guarded(4 * M, dev), Guarded(4 * M, dev), Guarded(4 * M, dev)
    rc = L.ab200_head_ce_forward(emb.data_ptr(), table.data_ptr(), tgt.data_ptr(), M, Z, 64, 0.2, lse.ptr(), tl.ptr(), None,
                                 None if dist is None else dist.data_ptr(), None if dist is None else ed.ptr(), ws.ptr(), ws.n, st)
    _lib.check(rc, "ab200_head_ce_forward")
    torch.cuda.synchronize()
    assert ws.intact() and lse.intact() and tl.intact() and ed.intact()
    gr, g2 = torch.rand(M, generator=g).to(dev), torch.rand(M, generator=g).to(dev)
    wb = Guarded(L.ab200_head_ce_backward_workspace_bytes(M, Z, 64), dev)
    de, dt = Guarded(4 * M * 64, dev), Guarded(4 * Z * 64, dev)
    rc = L.ab200_head_ce_backward(emb.data_ptr(), table.data_ptr(), tgt.data_ptr(), lse.ptr(), gr.data_ptr(),
                                  None if dist is None else g2.data_ptr(), None if dist is None else ed.ptr(),
                                  None if dist is None else dist.data_ptr(), M, Z, 64, 0.2, de.ptr(), dt.ptr(), wb.ptr(), wb.n, st)
    _lib.check(rc, "ab200_head_ce_backward")
    stt = C.c_int32(0)
    _lib.check(L.ab200_head_ce_backward_status(wb.ptr(), M, Z, C.byref(stt), st), "status")
    assert stt.value == 0
    assert wb.intact() and de.intact() and dt.intact() and lse.intact() and ed.intact()
    assert torch.isfinite(de.as_f32(M, 64)).all() and torch.isfinite(dt.as_f32(Z, 64)).all()


def test_sde_and_adam_kernels_stay_inside_their_buffers():
    from ananke_abm_b200 import _lib
    dev = _cuda()
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    B, D = 333, 64
    y, f = torch.randn(B, D, device=dev), torch.randn(B, D, device=dev)
    gd = torch.full((D,), 0.1, device=dev)
    out, xi = Guarded(4 * B * D, dev), Guarded(4 * B * D, dev)
    _lib.check(L.ab200_sde_euler_step(y.data_ptr(), f.data_ptr(), gd.data_ptr(), 0, B, D, 0.01, 7, 3, out.ptr(), xi.ptr(), st), "sde")
    n = 100_003
    p, m, v = Guarded(4 * n, dev), Guarded(4 * n, dev), Guarded(4 * n, dev)
    p.view.view(torch.float32).normal_(); m.view.zero_(); v.view.zero_()
    grad = torch.randn(n, device=dev)
    ss = Guarded(8, dev)
    _lib.check(L.ab200_grad_sumsq(grad.data_ptr(), n, ss.ptr(), st), "sumsq")
    _lib.check(L.ab200_adam_step(p.ptr(), grad.data_ptr(), m.ptr(), v.ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1.0, ss.ptr(), st), "adam")
    torch.cuda.synchronize()
    assert all(b.intact() for b in (out, xi, p, m, v, ss))


class _TorchProxy:
    """stands in for the `torch` module inside ananke_abm_b200.stage: CUDA `empty` / `zeros` come back as guarded views"""

    def __init__(self, real, log, dev):
        self._real, self._log, self._dev = real, log, dev

    def __getattr__(self, k):
        return getattr(self._real, k)

    def _guarded(self, size, kw, zero):
        real = self._real
        device = kw.get("device", None)
        if device is None or real.device(device).type != "cuda":
            return (real.zeros if zero else real.empty)(*size, **kw)
        shape = tuple(size[0]) if (len(size) == 1 and isinstance(size[0], (tuple, list, real.Size))) else tuple(int(s) for s in size)
        dtype = kw.get("dtype", real.float32)
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * real.empty((), dtype=dtype).element_size()
        gbuf = Guarded(max(nbytes, 0), self._dev)
        self._log.append(gbuf)
        t = gbuf.view.view(dtype).view(shape) if nbytes else real.empty(shape, dtype=dtype, device=device)
        if zero and nbytes:
            t.zero_()
        return t

    def empty(self, *size, **kw):
        return self._guarded(size, kw, False)

    def zeros(self, *size, **kw):
        return self._guarded(size, kw, True)


@pytest.mark.parametrize("B,method", [(300, "rk4"), (129, "dopri5"), (1, "dopri5"), (257, "rk4")])
def test_stage_path_stays_inside_its_buffers(B, method, monkeypatch):
    """The tensor-core training path (stage_fwd / stage_bwd / wgrad / elementwise kernels, weight image, spill and partial
    buffers) at ragged batch sizes: every buffer stage.py allocates is guarded; results stay finite."""
    import ananke_abm_b200 as ab
    from ananke_abm_b200 import stage
    dev = _cuda()
    log = []
    monkeypatch.setattr(stage, "torch", _TorchProxy(torch, log, dev))
    torch.manual_seed(1)
    m = ab.ModeSepModel(8, ab.ModeSepConfig()).to(dev)
    y0 = (torch.randn(B, 160, device=dev) * 0.3).requires_grad_(True)
    t = torch.linspace(0.0, 1.0, 4, device=dev)
    yp = ab.odeint(m.odefunc, y0, t, method=method, rtol=1e-4, atol=1e-4, options={"precision": "bf16"})
    yp[:, :, :128].square().mean().backward()
    torch.cuda.synchronize()
    assert len(log) > 10                                   # the proxy really was on the allocation path
    assert all(gb.intact() for gb in log)
    assert torch.isfinite(yp).all() and torch.isfinite(y0.grad).all()
    assert all(torch.isfinite(p.grad).all() for p in m.odefunc.parameters())
