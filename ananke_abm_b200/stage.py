"""Tensor-core STAGE path: explicit Runge-Kutta solvers assembled on the host from one-drift-evaluation launches
(`ab200_stage_forward` / `ab200_stage_backward` / `ab200_wgrad_*`, include/ananke_b200.h).

The drift of both reference models is second order (dp/dt = v), so every stage input, step solution, dense-output
row and embedded error of torchdiffeq's solvers is linear in the step's base state y0 = [p0, v0, h] and the stage
accelerations a_j; the host only computes those coefficients (`Tableau.combo`) -- all arithmetic runs in the library.

  rk4     torchdiffeq fixed_grid.py RK4 / rk_common.py rk4_alt_step_func (3/8 rule), one step per grid interval
          -- the call at /root/reference/src/ananke_abm/models/mode_sep/architecture/model.py:184-191
  dopri5  torchdiffeq dopri5.py + rk_common.py RKAdaptiveStepsizeODESolver (FSAL, dense output)
          -- the call at /root/reference/src/ananke_abm/models/latent_ode/architecture/model.py:196
The backward pass is the discrete adjoint of the accepted steps == reverse-mode autograd through the solver ops,
which is what the reference's training loops compute (mode_sep/train/train.py:162, latent_ode/train/train.py:73).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .drift import DriftSpec

MAX_A = 7
TM = 128
SAVE_LEVELS = {"none": 0, "inputs": 1, "all": 2}
FWD_FORMATS = {"bf16": 0, "fp16": 1, "fp16x2": 2}


def fwd_format_code(v) -> int:
    """operand format of the forward stage kernels from a name, a code, or the legacy fp16_forward bool"""
    if isinstance(v, str):
        if v not in FWD_FORMATS:
            raise ValueError(f"unknown forward operand format {v!r}; choose from {sorted(FWD_FORMATS)}")
        return FWD_FORMATS[v]
    if isinstance(v, bool):
        return 1 if v else 0
    v = int(v)
    if v not in (0, 1, 2):
        raise ValueError(f"unknown forward operand format code {v}")
    return v


class StageDesc(C.Structure):
    """Mirror of `ab200_stage_desc`."""
    _fields_ = [
        ("n_a", C.c_int32),
        ("in_cpv", C.c_float), ("in_cpa", C.c_float * MAX_A), ("in_cva", C.c_float * MAX_A),
        ("t", C.c_float),
        ("out_cpv", C.c_float), ("out_cpa", C.c_float * (MAX_A + 1)), ("out_cva", C.c_float * (MAX_A + 1)),
        ("err_pa", C.c_float * (MAX_A + 1)), ("err_va", C.c_float * (MAX_A + 1)),
        ("rtol", C.c_float), ("atol", C.c_float),
    ]


@dataclass
class Combo:
    """p = p0 + cpv v0 + sum cpa[j] a_j ;  v = v0 + sum cva[j] a_j"""
    cpv: float
    cpa: List[float]
    cva: List[float]


class Tableau:
    def __init__(self, c: Sequence[float], beta: Sequence[Sequence[float]], b: Sequence[float],
                 b_err: Optional[Sequence[float]] = None, b_mid: Optional[Sequence[float]] = None):
        self.c, self.beta, self.b, self.b_err, self.b_mid = list(c), [list(r) for r in beta], list(b), b_err, b_mid
        self.s = len(self.beta)

    def combo(self, w: Sequence[float], dt: float) -> Combo:
        """y0 + dt sum_j w_j k_j with k_j = (v_in_j, a_j) written over (p0, v0, a_1..a_n)."""
        n = len(w)
        cva_rows = [[dt * x for x in self.beta[j]] + [0.0] * (n - len(self.beta[j])) for j in range(n)]
        cpa = [dt * sum(w[j] * cva_rows[j][l] for j in range(n)) for l in range(n)]
        return Combo(dt * sum(w), cpa, [dt * x for x in w])

    def stage_input(self, i: int, dt: float) -> Combo:
        """input of stage i (0-based): combination of a_0..a_{i-1}"""
        return self.combo(self.beta[i], dt)


RK38 = Tableau(c=[0.0, 1 / 3, 2 / 3, 1.0], beta=[[], [1 / 3], [-1 / 3, 1.0], [1.0, -1.0, 1.0]], b=[1 / 8, 3 / 8, 3 / 8, 1 / 8])

_DP_BETA = [
    [],
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_DP_C_SOL = [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0]
_DP_C_ERR = [35 / 384 - 1951 / 21600, 0.0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
             -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0]
_DP_C_MID = [6025192743 / 30085553152 / 2, 0.0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
             187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]
DOPRI5 = Tableau(c=[0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0], beta=_DP_BETA, b=_DP_C_SOL, b_err=_DP_C_ERR, b_mid=_DP_C_MID)


def dopri5_interp_weights(x: float) -> List[float]:
    """Weights W_j(x) with y(t0 + x dt) = y0 + dt sum_j W_j k_j: torchdiffeq's quartic dense output
    (interp.py `_interp_fit` / `_interp_evaluate`) expanded over the seven stage derivatives."""
    e1 = [1.0, 0, 0, 0, 0, 0, 0]
    e7 = [0, 0, 0, 0, 0, 0, 1.0]
    cs, cm = _DP_C_SOL, _DP_C_MID
    out = []
    for j in range(7):
        c2 = e7[j] - 4 * e1[j] - 5 * cs[j] + 16 * cm[j]
        c3 = 5 * e1[j] - 3 * e7[j] + 14 * cs[j] - 32 * cm[j]
        c4 = 2 * (e7[j] - e1[j]) - 8 * cs[j] + 16 * cm[j]
        out.append(x * e1[j] + x * x * c2 + x ** 3 * c3 + x ** 4 * c4)
    return out


def _ptr_array(tensors: Sequence[torch.Tensor]):
    n = max(len(tensors), 1)
    return (C.c_void_p * n)(*[t.data_ptr() for t in tensors])


def _fill(dst, vals):
    for i, v in enumerate(vals):
        dst[i] = float(v)


class TcEngine:
    """Holds the packed bf16 weight image for one parameter version and issues stage launches on the current stream."""

    def __init__(self, spec: DriftSpec, w_flat: torch.Tensor):
        self.L = _lib.lib()
        self.spec, self.desc = spec, spec.desc
        self.dev = w_flat.device
        nbytes = self.L.ab200_stage_image_bytes(C.byref(self.desc))
        if nbytes == 0:
            raise _lib.Ab200Error("the tensor-core stage path is instantiated for the mode_sep drift shape only")
        self.image = torch.empty(int(nbytes), dtype=torch.uint8, device=self.dev)
        self.w = w_flat.detach().contiguous().float()
        _lib.check(self.L.ab200_stage_pack(C.byref(self.desc), self.w.data_ptr(), self.image.data_ptr(), self.image.numel(),
                                           _stream()), "ab200_stage_pack")
        off_i, off_p = C.c_int64(0), C.c_int64(0)
        self.L.ab200_stage_status_offset(C.byref(self.desc), C.byref(off_i), C.byref(off_p))
        self._img_status, self._part_status = int(off_i.value), int(off_p.value)
        self.P, self.H = self.desc.pos_dim, self.desc.ctx_dim
        self.D = 2 * self.P + self.H
        # forward stage evaluations: operand format of ab200_stage_forward* (FWD_FORMATS).  "fp16" = IEEE fp16 operands
        # (values of this net stay far inside the fp16 range; 8x less rounding noise than bf16 at the same speed);
        # "fp16x2" = fp16 weights, activations as two-term fp16 splits: no activation rounding at all, which is what the
        # embedded error estimate of an adaptive solver needs (dopri5_forward selects it).  Backward kernels keep bf16.
        self.fwd_format = FWD_FORMATS["fp16"]

    # ---- forward ------------------------------------------------------------------------------------
    def stage_forward(self, y0, a: Sequence[torch.Tensor], cin: Combo, t: float, B: int, a_out=None, y_out=None, cout: Optional[Combo] = None,
                      err_sumsq=None, cerr: Optional[Combo] = None, rtol: float = 0.0, atol: float = 0.0, fp16: Optional[bool] = None) -> None:
        s = StageDesc()
        s.n_a = len(a)
        s.in_cpv = cin.cpv
        _fill(s.in_cpa, cin.cpa[:len(a)])
        _fill(s.in_cva, cin.cva[:len(a)])
        s.t = float(t)
        if cout is not None:
            s.out_cpv = cout.cpv
            _fill(s.out_cpa, cout.cpa[:len(a) + 1])
            _fill(s.out_cva, cout.cva[:len(a) + 1])
        if cerr is not None:
            _fill(s.err_pa, cerr.cpa[:len(a) + 1])
            _fill(s.err_va, cerr.cva[:len(a) + 1])
        s.rtol, s.atol = float(rtol), float(atol)
        rc = self.L.ab200_stage_forward(C.byref(self.desc), self.image.data_ptr(), y0.data_ptr(), C.cast(_ptr_array(a), C.c_void_p),
                                        C.byref(s), B, None if a_out is None else a_out.data_ptr(),
                                        None if y_out is None else y_out.data_ptr(),
                                        None if err_sumsq is None else err_sumsq.data_ptr(),
                                        self.fwd_format if fp16 is None else fwd_format_code(fp16), _stream())
        _lib.check(rc, "ab200_stage_forward")

    def stage_forward_fused(self, y0, a_bufs: Sequence[torch.Tensor], stages: Sequence, B: int, y_out=None, cout: Optional[Combo] = None,
                            err_sumsq=None, cerr: Optional[Combo] = None, rtol: float = 0.0, atol: float = 0.0,
                            fp16: Optional[bool] = None, x_outs: Optional[Sequence[int]] = None, save_level: int = 0) -> None:
        """Consecutive stages of one step in ONE launch.  `stages` = [(n_a, Combo in, t, a_out tensor or None), ...]: stage s
        reads a_bufs[:n_a] (possibly written by earlier stages of this call) and writes its a_out; the last stage may
        also produce y_out (cout) and the error norm (cerr)."""
        n = len(stages)
        descs = (StageDesc * n)()
        outs = (C.c_void_p * n)()
        for i, (n_a, cin, t, a_out) in enumerate(stages):
            s = descs[i]
            s.n_a = n_a
            s.in_cpv = cin.cpv
            _fill(s.in_cpa, cin.cpa[:n_a])
            _fill(s.in_cva, cin.cva[:n_a])
            s.t = float(t)
            outs[i] = None if a_out is None else a_out.data_ptr()
        last = descs[n - 1]
        n_last = stages[-1][0]
        if cout is not None:
            last.out_cpv = cout.cpv
            _fill(last.out_cpa, cout.cpa[:n_last + 1])
            _fill(last.out_cva, cout.cva[:n_last + 1])
        if cerr is not None:
            _fill(last.err_pa, cerr.cpa[:n_last + 1])
            _fill(last.err_va, cerr.cva[:n_last + 1])
        last.rtol, last.atol = float(rtol), float(atol)
        ptrs = (C.c_void_p * MAX_A)(*([t.data_ptr() for t in a_bufs[:MAX_A]] + [None] * (MAX_A - min(len(a_bufs), MAX_A))))
        if x_outs is not None:      # split-activation format; every stage saves its backward operands (device pointers, one per stage)
            assert len(x_outs) == n and save_level in (1, 2)
            rc = self.L.ab200_stage_forward_fused_save(C.byref(self.desc), self.image.data_ptr(), y0.data_ptr(), C.cast(ptrs, C.c_void_p),
                                                       C.cast(descs, C.c_void_p), n, C.cast(outs, C.c_void_p), B,
                                                       None if y_out is None else y_out.data_ptr(),
                                                       None if err_sumsq is None else err_sumsq.data_ptr(),
                                                       C.cast((C.c_void_p * n)(*[int(x) for x in x_outs]), C.c_void_p), int(save_level),
                                                       _stream())
            _lib.check(rc, "ab200_stage_forward_fused_save")
            return
        rc = self.L.ab200_stage_forward_fused(C.byref(self.desc), self.image.data_ptr(), y0.data_ptr(), C.cast(ptrs, C.c_void_p),
                                              C.cast(descs, C.c_void_p), n, C.cast(outs, C.c_void_p), B,
                                              None if y_out is None else y_out.data_ptr(),
                                              None if err_sumsq is None else err_sumsq.data_ptr(),
                                              self.fwd_format if fp16 is None else fwd_format_code(fp16), _stream())
        _lib.check(rc, "ab200_stage_forward_fused")

    def combine(self, y0, a: Sequence[torch.Tensor], c: Combo, B: int, out) -> None:
        n = len(a)
        cpa = (C.c_float * max(n, 1))(*[float(x) for x in c.cpa[:n]])
        cva = (C.c_float * max(n, 1))(*[float(x) for x in c.cva[:n]])
        rc = self.L.ab200_pv_combine(C.byref(self.desc), y0.data_ptr(), C.cast(_ptr_array(a), C.c_void_p), n, float(c.cpv),
                                     C.cast(cpa, C.c_void_p), C.cast(cva, C.c_void_p), B, out.data_ptr(), _stream())
        _lib.check(rc, "ab200_pv_combine")

    def combine_rowmajor(self, y0, a: Sequence[torch.Tensor], c: Combo, B: int, out_rowmajor) -> None:
        """`combine` written straight into a row-major [B, D] tensor (a trajectory row)."""
        n = len(a)
        assert out_rowmajor.is_contiguous()
        cpa = (C.c_float * max(n, 1))(*[float(x) for x in c.cpa[:n]])
        cva = (C.c_float * max(n, 1))(*[float(x) for x in c.cva[:n]])
        rc = self.L.ab200_pv_combine_rowmajor(C.byref(self.desc), y0.data_ptr(), C.cast(_ptr_array(a), C.c_void_p), n, float(c.cpv),
                                              C.cast(cpa, C.c_void_p), C.cast(cva, C.c_void_p), B, out_rowmajor.data_ptr(), _stream())
        _lib.check(rc, "ab200_pv_combine_rowmajor")

    def dopri5_attempt(self, y0, A: Sequence[torch.Tensor], t0: float, dt: float, B: int, y_out, err_sumsq, rtol: float, atol: float,
                       x_blobs=None, save_level: int = 0) -> None:
        """one attempted Dormand-Prince step in ONE call: stages 2..7 fused, descriptors built in C (ab200_dopri5_attempt);
        `x_blobs` (6 * xblob_bytes(B, save_level) bytes, split-activation format only) receives what the backward pass would
        otherwise recompute (level 1: stage inputs; level 2: + hidden activations and ReLU masks) as its operand images"""
        ptrs = (C.c_void_p * 7)(*[t.data_ptr() for t in A[:7]])
        rc = self.L.ab200_dopri5_attempt(C.byref(self.desc), self.image.data_ptr(), y0.data_ptr(), C.cast(ptrs, C.c_void_p), float(t0),
                                         float(dt), B, y_out.data_ptr(), err_sumsq.data_ptr(), float(rtol), float(atol), self.fwd_format,
                                         None if x_blobs is None else x_blobs.data_ptr(), int(save_level), _stream())
        _lib.check(rc, "ab200_dopri5_attempt")

    def xblob_bytes(self, B: int, save_level: int) -> int:
        return int(self.L.ab200_stage_xblob_bytes(C.byref(self.desc), B, int(save_level)))

    def dopri5_dense_rows(self, y0, A: Sequence[torch.Tensor], dt: float, xs: Sequence[float], B: int, outs: Sequence[torch.Tensor]) -> None:
        """dense-output rows of an accepted step at relative positions xs, one pass (ab200_dopri5_dense_rows)"""
        n = len(xs)
        ptrs = (C.c_void_p * 7)(*[t.data_ptr() for t in A[:7]])
        xa = (C.c_double * n)(*[float(x) for x in xs])
        op = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
        rc = self.L.ab200_dopri5_dense_rows(C.byref(self.desc), y0.data_ptr(), C.cast(ptrs, C.c_void_p), float(dt), n,
                                            C.cast(xa, C.c_void_p), B, C.cast(op, C.c_void_p), _stream())
        _lib.check(rc, "ab200_dopri5_dense_rows")

    def combine_rowmajor_multi(self, y0, a: Sequence[torch.Tensor], combos: Sequence[Combo], B: int, outs: Sequence[torch.Tensor]) -> None:
        """several dense-output rows of the same step in ONE pass (y0 and the accelerations are read once)"""
        n, nr = len(a), len(combos)
        cpv = (C.c_float * nr)(*[float(c.cpv) for c in combos])
        cpa = (C.c_float * (nr * 8))()
        cva = (C.c_float * (nr * 8))()
        for q, c in enumerate(combos):
            for j in range(n):
                cpa[q * 8 + j] = float(c.cpa[j])
                cva[q * 8 + j] = float(c.cva[j])
        op = (C.c_void_p * nr)(*[o.data_ptr() for o in outs])
        assert all(o.is_contiguous() for o in outs)
        rc = self.L.ab200_pv_combine_rowmajor_multi(C.byref(self.desc), y0.data_ptr(), C.cast(_ptr_array(a), C.c_void_p), n, nr,
                                                    C.cast(cpv, C.c_void_p), C.cast(cpa, C.c_void_p), C.cast(cva, C.c_void_p), B,
                                                    C.cast(op, C.c_void_p), _stream())
        _lib.check(rc, "ab200_pv_combine_rowmajor_multi")

    # ---- backward -----------------------------------------------------------------------------------
    def backward_begin(self, B: int, stages_per_flush: int) -> None:
        self.ntiles = (B + TM - 1) // TM
        self.nblobs = self.ntiles * stages_per_flush
        nb = self.L.ab200_stage_spill_bytes(C.byref(self.desc), self.nblobs)
        self.spill = torch.empty(int(nb), dtype=torch.uint8, device=self.dev)
        npart = self.L.ab200_wgrad_partial_bytes(C.byref(self.desc))
        self.partial = torch.zeros(int(npart), dtype=torch.uint8, device=self.dev)
        self.used = 0
        self.x_ring = []          # per filled run of ntiles blobs: device pointer of a forward-saved buffer, or None
        self.x_level = 0          # save level of those buffers (one level per backward pass)

    def combine_backward(self, g, c: Combo, B: int, G_y0, G_a: Sequence[torch.Tensor], accumulate: bool) -> None:
        n = len(G_a)
        cpa = (C.c_float * max(n, 1))(*[float(x) for x in c.cpa[:n]])
        cva = (C.c_float * max(n, 1))(*[float(x) for x in c.cva[:n]])
        rc = self.L.ab200_pv_combine_backward(C.byref(self.desc), g.data_ptr(), n, float(c.cpv), C.cast(cpa, C.c_void_p),
                                              C.cast(cva, C.c_void_p), B, G_y0.data_ptr(), C.cast(_ptr_array(G_a), C.c_void_p),
                                              1 if accumulate else 0, _stream())
        _lib.check(rc, "ab200_pv_combine_backward")

    def stage_backward(self, y0, a: Sequence[torch.Tensor], cin: Combo, t: float, B: int, g_base, gx: Sequence[torch.Tensor],
                       dp: Sequence[float], dv: Sequence[float], gx_out) -> None:
        """dL/da_out = g_base + sum_l dp[l] gx[l].p + dv[l] gx[l].v  ->  gx_out = dL/d(stage input) (+ blobs)."""
        if self.used + self.ntiles > self.nblobs:
            self.flush()
        s = StageDesc()
        s.n_a = len(a)
        s.in_cpv = cin.cpv
        _fill(s.in_cpa, cin.cpa[:len(a)])
        _fill(s.in_cva, cin.cva[:len(a)])
        s.t = float(t)
        n = len(gx)
        dpa = (C.c_float * max(n, 1))(*[float(x) for x in dp])
        dva = (C.c_float * max(n, 1))(*[float(x) for x in dv])
        rc = self.L.ab200_stage_backward(C.byref(self.desc), self.image.data_ptr(), y0.data_ptr(), C.cast(_ptr_array(a), C.c_void_p),
                                         C.byref(s), B, None if g_base is None else g_base.data_ptr(),
                                         C.cast(_ptr_array(gx), C.c_void_p), n, C.cast(dpa, C.c_void_p), C.cast(dva, C.c_void_p),
                                         gx_out.data_ptr(), self.spill.data_ptr(), self.spill.numel(), self.used, self.nblobs,
                                         self.partial.data_ptr(), _stream())
        _lib.check(rc, "ab200_stage_backward")
        self.used += self.ntiles
        self.x_ring.append(None)

    def combine_backward_multi(self, sources: Sequence, B: int, G_y0, G_a: Sequence[torch.Tensor], accumulate: bool, add_a=None,
                               add_index: int = 0) -> None:
        """`sources` = [(g, Combo), ...]: all folded into G_y0 / G_a in passes of up to 6 sources.  g is a blocked tensor, or a
        contiguous ROW-MAJOR [B, D] row of the caller's gradient tensor (2-D: read in place, no transposed copy).
        `add_a` (blocked [Bp, P]) is added to G_a[add_index] in the first pass."""
        n_a = len(G_a)
        for s0 in range(0, len(sources), 6):
            grp = sources[s0:s0 + 6]
            n = len(grp)
            rm_mask = 0
            for i, (g, _) in enumerate(grp):
                if g.dim() == 2:
                    assert g.is_contiguous() and g.shape == (B, self.D) and g.dtype == torch.float32
                    rm_mask |= 1 << i
            gp = (C.c_void_p * n)(*[g.data_ptr() for g, _ in grp])
            cpv = (C.c_float * n)(*[float(c.cpv) for _, c in grp])
            cpa = (C.c_float * (n * 8))()
            cva = (C.c_float * (n * 8))()
            for i, (_, c) in enumerate(grp):
                for j in range(n_a):
                    cpa[i * 8 + j] = float(c.cpa[j])
                    cva[i * 8 + j] = float(c.cva[j])
            rc = self.L.ab200_pv_combine_backward_multi(C.byref(self.desc), C.cast(gp, C.c_void_p), n, C.cast(cpv, C.c_void_p),
                                                        C.cast(cpa, C.c_void_p), C.cast(cva, C.c_void_p), n_a, B, G_y0.data_ptr(),
                                                        C.cast(_ptr_array(G_a), C.c_void_p), 1 if (accumulate or s0 > 0) else 0, rm_mask,
                                                        add_a.data_ptr() if (add_a is not None and s0 == 0) else None, int(add_index), _stream())
            _lib.check(rc, "ab200_pv_combine_backward_multi")

    def stage_upstream(self, g_base, gx: Sequence[torch.Tensor], dp: Sequence[float], dv: Sequence[float], B: int, out) -> None:
        """out = g_base + sum_l dp[l] gx[l].p + dv[l] gx[l].v (the upstream gradient of a stage, as a buffer)."""
        n = len(gx)
        dpa = (C.c_float * max(n, 1))(*[float(x) for x in dp])
        dva = (C.c_float * max(n, 1))(*[float(x) for x in dv])
        rc = self.L.ab200_stage_upstream(C.byref(self.desc), g_base.data_ptr(), C.cast(_ptr_array(gx), C.c_void_p), n,
                                         C.cast(dpa, C.c_void_p), C.cast(dva, C.c_void_p), B, out.data_ptr(), _stream())
        _lib.check(rc, "ab200_stage_upstream")

    def stage_backward_fused(self, y0, a_bufs: Sequence[torch.Tensor], stages: Sequence, B: int, x_blobs: Optional[Sequence] = None,
                             save_level: int = 0, y0_accum=None, upstream=None) -> None:
        """The backward stages of one step in ONE launch.  `stages` (latest stage first) =
        [(n_a, Combo in, t, g_base tensor or None, [(src, dp, dv), ...], gx_out tensor)], where `src` is the position of an
        earlier entry of `stages` whose gx_out feeds this stage's upstream gradient.  `x_blobs[i]`: device pointer of the
        forward-saved buffer of entry i (`dopri5_attempt`, same `save_level`), or None to rebuild / recompute from (y0, a_j).
        `y0_accum` (blocked [Bp, D], initialised with the step-level dL/dy0) and / or `upstream` = (g_base, [(dp, dv)] per entry
        of `stages`, out) add one GATHER entry to the launch: it folds every stage's gx into dL/dy0 (in place, no `adjoint_gather`
        pass) and writes g_base + sum dp gx.p + dv gx.v to `out` (the gradient of the previous step's FSAL evaluation)."""
        n_real = len(stages)
        if upstream is not None or y0_accum is not None:
            coef = upstream[1] if upstream is not None else [(0.0, 0.0)] * n_real
            stages = list(stages) + [(0, Combo(0.0, [], []), 0.0, upstream[0] if upstream is not None else None,
                                      [(i, dp_, dv_) for i, (dp_, dv_) in enumerate(coef)], None)]
        n = len(stages)
        xs = (list(x_blobs) if x_blobs is not None else [None] * n_real) + [None] * (n - n_real)
        assert len(xs) == n
        if any(x is not None for x in xs):
            assert self.x_level in (0, save_level), "one save level per backward pass"
            self.x_level = save_level
        if self.used + n_real * self.ntiles > self.nblobs:
            self.flush()
        descs = (StageDesc * n)()
        g_base = (C.c_void_p * n)()
        gx_out = (C.c_void_p * n)()
        up_out = (C.c_void_p * n)()
        if upstream is not None:
            up_out[n - 1] = upstream[2].data_ptr()
        n_g = (C.c_int32 * n)()
        src = (C.c_int32 * (n * MAX_A))()
        dp = (C.c_float * (n * MAX_A))()
        dv = (C.c_float * (n * MAX_A))()
        for i, (n_a, cin, t, gb, sources, gout) in enumerate(stages):
            s = descs[i]
            s.n_a = n_a
            s.in_cpv = cin.cpv
            _fill(s.in_cpa, cin.cpa[:n_a])
            _fill(s.in_cva, cin.cva[:n_a])
            s.t = float(t)
            g_base[i] = None if gb is None else gb.data_ptr()
            gx_out[i] = None if gout is None else gout.data_ptr()
            n_g[i] = len(sources)
            for l, (sidx, a_, b_) in enumerate(sources):
                src[i * MAX_A + l] = sidx
                dp[i * MAX_A + l] = float(a_)
                dv[i * MAX_A + l] = float(b_)
        ptrs = (C.c_void_p * MAX_A)(*([t.data_ptr() for t in a_bufs[:MAX_A]] + [None] * (MAX_A - min(len(a_bufs), MAX_A))))
        rc = self.L.ab200_stage_backward_fused(C.byref(self.desc), self.image.data_ptr(), y0.data_ptr(), C.cast(ptrs, C.c_void_p),
                                               C.cast(descs, C.c_void_p), n, C.cast(g_base, C.c_void_p), C.cast(gx_out, C.c_void_p),
                                               C.cast(n_g, C.c_void_p), C.cast(src, C.c_void_p), None, C.cast(dp, C.c_void_p),
                                               C.cast(dv, C.c_void_p), B, self.spill.data_ptr(), self.spill.numel(), self.used,
                                               self.nblobs, self.partial.data_ptr(),
                                               C.cast((C.c_void_p * n)(*xs), C.c_void_p) if any(x is not None for x in xs) else None,
                                               int(save_level), None if y0_accum is None else y0_accum.data_ptr(),
                                               C.cast(up_out, C.c_void_p) if upstream is not None else None, _stream())
        _lib.check(rc, "ab200_stage_backward_fused")
        self.used += n_real * self.ntiles
        self.x_ring.extend(xs[:n_real])

    def adjoint_gather_upstream(self, base, gx: Sequence[torch.Tensor], cpv: Sequence[float], B: int, out, g_base,
                                dp: Sequence[float], dv: Sequence[float], g_a_out) -> None:
        """adjoint_gather + stage_upstream over the same gx list in one pass."""
        n = len(gx)
        ca = (C.c_float * n)(*[float(x) for x in cpv])
        dpa = (C.c_float * n)(*[float(x) for x in dp])
        dva = (C.c_float * n)(*[float(x) for x in dv])
        rc = self.L.ab200_adjoint_gather_upstream(C.byref(self.desc), base.data_ptr(), C.cast(_ptr_array(gx), C.c_void_p), n,
                                                  C.cast(ca, C.c_void_p), B, out.data_ptr(), g_base.data_ptr(), C.cast(dpa, C.c_void_p),
                                                  C.cast(dva, C.c_void_p), g_a_out.data_ptr(), _stream())
        _lib.check(rc, "ab200_adjoint_gather_upstream")

    def adjoint_gather(self, base, gx: Sequence[torch.Tensor], cpv: Sequence[float], B: int, out) -> None:
        n = len(gx)
        ca = (C.c_float * max(n, 1))(*[float(x) for x in cpv])
        rc = self.L.ab200_adjoint_gather(C.byref(self.desc), base.data_ptr(), C.cast(_ptr_array(gx), C.c_void_p), n,
                                         C.cast(ca, C.c_void_p), B, out.data_ptr(), _stream())
        _lib.check(rc, "ab200_adjoint_gather")

    def flush(self) -> None:
        if self.used:
            nx = len(self.x_ring) if any(x is not None for x in self.x_ring) else 0
            rc = self.L.ab200_wgrad_accumulate(C.byref(self.desc), self.spill.data_ptr(), self.nblobs, self.used,
                                               self.partial.data_ptr(),
                                               C.cast((C.c_void_p * nx)(*self.x_ring), C.c_void_p) if nx else None, nx, self.ntiles,
                                               self.x_level, _stream())
            _lib.check(rc, "ab200_wgrad_accumulate")
            self.used = 0
            self.x_ring = []

    def backward_end(self) -> torch.Tensor:
        self.flush()
        gw = torch.empty_like(self.w)
        _lib.check(self.L.ab200_wgrad_finalize(C.byref(self.desc), self.partial.data_ptr(), gw.data_ptr(), _stream()),
                   "ab200_wgrad_finalize")
        self.spill = None
        self.check_status()          # one host read per backward pass: a timed-out tensor-core kernel must not yield gradients
        return gw

    def check_status(self) -> None:
        """Host sync: raises if any tensor-core kernel that used this engine's image / partial buffer hit its bounded
        barrier wait (its results are garbage).  Called by the product path at the end of every rk4 forward, of every
        backward pass and when dopri5 sees a poisoned error norm."""
        st = int(self.image[self._img_status:self._img_status + 4].view(torch.int32).item())
        sp = 0
        if getattr(self, "partial", None) is not None:
            sp = int(self.partial[self._part_status:self._part_status + 4].view(torch.int32).item())
        if st or sp:
            raise _lib.Ab200Error(f"tensor-core kernel barrier timeout (stage status {st}, wgrad status {sp})")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


_SIDE_STREAMS: dict = {}


def _side_stream(device) -> "torch.cuda.Stream":
    """one extra stream per device for the per-attempt read of dopri5's error norm (see dopri5_forward)"""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=key)
    return _SIDE_STREAMS[key]


# --------------------------------------------------------------------------------------------------------
# blocked buffers
# --------------------------------------------------------------------------------------------------------
def padded_rows(B: int) -> int:
    return (B + TM - 1) // TM * TM


def blocked_zeros(B: int, F: int, device) -> torch.Tensor:
    """A zeroed tile-blocked [Bp, F] buffer (flat storage; see include/ananke_b200.h for the layout)."""
    return torch.zeros(padded_rows(B) * F, dtype=torch.float32, device=device)


def blocked_empty(B: int, F: int, device) -> torch.Tensor:
    """An uninitialised tile-blocked [Bp, F] buffer for kernel OUTPUTS (ab200_stage_forward* write every row of a_out / y_out,
    zeros in the padding rows of the last tile)."""
    return torch.empty(padded_rows(B) * F, dtype=torch.float32, device=device)


def rows_block(src: torch.Tensor, dst: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """row-major [B, F] -> blocked (dst = src, or dst += src)."""
    L = _lib.lib()
    B, F = src.shape
    src = src.contiguous()
    if dst is None:
        dst = torch.empty(padded_rows(B) * F, dtype=torch.float32, device=src.device)
        assert not accumulate
    _lib.check(L.ab200_rows_block(src.data_ptr(), dst.data_ptr(), B, F, 1 if accumulate else 0, _stream()), "ab200_rows_block")
    return dst


def rows_unblock(src: torch.Tensor, B: int, F: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """blocked -> row-major [B, F]."""
    L = _lib.lib()
    if out is None:
        out = torch.empty((B, F), dtype=torch.float32, device=src.device)
    assert out.is_contiguous()
    _lib.check(L.ab200_rows_unblock(src.data_ptr(), out.data_ptr(), B, F, _stream()), "ab200_rows_unblock")
    return out


# --------------------------------------------------------------------------------------------------------
# fixed-grid rk4 (3/8 rule) with the discrete adjoint
# --------------------------------------------------------------------------------------------------------
def rk4_forward(eng: TcEngine, y0: torch.Tensor, t_host: Sequence[float], save_stages: bool, saved_operands: str = "none"):
    """y0 row-major [B, D] -> y_path row-major [T, B, D]; when `save_stages`, also the blocked per-step states
    yb [T, Bp*D] and stage accelerations acc [T-1, 3, Bp*P] the adjoint needs.  `saved_operands` ("none" = default | "inputs" | "all",
    training only): the four stage evaluations of every step run in the split-activation format and write what their backward pass would
    recompute (`ab200_stage_forward_fused_save`: the stage inputs, 352 B per agent-stage, or also the hidden activations and ReLU
    masks, 1,712 B) -- returned as xs [T-1, 4 * xblob_bytes]."""
    B, T = y0.shape[0], len(t_host)
    dev = y0.device
    Bp = padded_rows(B)
    level = SAVE_LEVELS[saved_operands] if save_stages else 0
    y_path = torch.empty((T, B, eng.D), dtype=torch.float32, device=dev)
    y_path[0].copy_(y0)
    n_keep = T - 1 if save_stages else 1
    acc = torch.zeros((n_keep, 3, Bp * eng.P), dtype=torch.float32, device=dev)
    yb = torch.zeros((T if save_stages else 2, Bp * eng.D), dtype=torch.float32, device=dev)
    per = eng.xblob_bytes(B, level) if level else 0
    xs = torch.empty((max(T - 1, 1), 4 * per), dtype=torch.uint8, device=dev) if level else None
    rows_block(y0, yb[0])
    for n in range(T - 1):
        t0, dt = float(t_host[n]), float(t_host[n + 1]) - float(t_host[n])
        A = acc[n if save_stages else 0]
        yn = yb[n if save_stages else n % 2]
        yn1 = yb[n + 1 if save_stages else (n + 1) % 2]
        stages = [(i, RK38.stage_input(i, dt), t0 + RK38.c[i] * dt, A[i]) for i in range(3)]
        stages.append((3, RK38.stage_input(3, dt), float(t_host[n + 1]), None))
        if level:
            base = xs[n].data_ptr()
            eng.stage_forward_fused(yn, [A[0], A[1], A[2]], stages, B, y_out=yn1, cout=RK38.combo(RK38.b, dt),
                                    x_outs=[base + i * per for i in range(4)], save_level=level)
        else:
            eng.stage_forward_fused(yn, [A[0], A[1], A[2]], stages, B, y_out=yn1, cout=RK38.combo(RK38.b, dt))
        rows_unblock(yn1, B, eng.D, out=y_path[n + 1])
    eng.check_status()
    return y_path, ((yb, acc, xs, level) if save_stages else None)


def stages_backward(eng: TcEngine, tab: Tableau, B: int, yn, A: Sequence[torch.Tensor], stage_times: Sequence[float], dt: float,
                    G_a_base: Sequence[Optional[torch.Tensor]], gx: Sequence[torch.Tensor], first: int, last: int, x_blobs=None,
                    save_level: int = 0, y0_accum=None, fsal_out=None, x_first: int = 1):
    """Backward of stages last..first of ONE explicit Runge-Kutta step in a single fused launch (latest stage first).
    Returns the stage input combinations (their cpv feed `adjoint_gather`).  `x_blobs`: what the step's forward attempt saved
    (tensor, stage i >= 1 at byte offset (i - 1) * xblob_bytes(B, save_level)) or None.  `y0_accum`: see `stage_backward_fused`.
    `fsal_out` (first >= 1): receives dL/da_1 = G_a_base[0] + the contributions of stages first..last -- the gradient w.r.t. the
    evaluation this step took over from the previous one (FSAL)."""
    combos = [tab.stage_input(i, dt) for i in range(last + 1)]
    order = list(range(last, first - 1, -1))
    pos = {i: k for k, i in enumerate(order)}
    stages = []
    for i in order:
        later = [l for l in range(i + 1, last + 1) if combos[l].cpa[i] != 0.0 or combos[l].cva[i] != 0.0]
        stages.append((i, combos[i], stage_times[i], G_a_base[i], [(pos[l], combos[l].cpa[i], combos[l].cva[i]) for l in later], gx[i]))
    xs = None
    if x_blobs is not None:
        per = eng.xblob_bytes(B, save_level)
        xs = [x_blobs.data_ptr() + (i - x_first) * per if i >= x_first else None for i in order]      # dopri5: stage 1 is the FSAL evaluation
    upstream = None
    if fsal_out is not None:
        assert first >= 1
        upstream = (G_a_base[0], [(combos[l].cpa[0], combos[l].cva[0]) for l in order], fsal_out)
    eng.stage_backward_fused(yn, list(A), stages, B, xs, save_level, y0_accum, upstream)
    return combos


def step_backward(eng: TcEngine, tab: Tableau, B: int, yn, A: Sequence[torch.Tensor], stage_times: Sequence[float], dt: float,
                  G_y0_base, G_a_base: Sequence[Optional[torch.Tensor]], gx: Sequence[torch.Tensor], out) -> None:
    """Adjoint of the stages of ONE explicit Runge-Kutta step (any tableau).  G_y0_base / G_a_base hold the step-level
    gradients w.r.t. y0 and the stage accelerations (from the step's linear outputs); `out` receives dL/dy0 of the step.
    `gx[i]` is scratch for stage i's dL/d(stage input)."""
    s = len(stage_times)
    combos = stages_backward(eng, tab, B, yn, A, stage_times, dt, G_a_base, gx, 0, s - 1)
    eng.adjoint_gather(G_y0_base, [gx[i] for i in range(s)], [combos[i].cpv for i in range(s)], B, out)


def rk4_backward(eng: TcEngine, t_host: Sequence[float], saved, grad_y_path: torch.Tensor):
    """-> (grad_y0 row-major [B, D], grad_w_flat).  Per step: ONE elementwise pass (the step-level gradients from dL/dy_{n+1} of the
    later steps and the caller's row-major dL/dy_path[n + 1], read in place), one fused launch of the four backward stages whose
    gather entry folds the stages' gx into dL/dy_n in place, one weight-gradient pass -- the same three launches as a dopri5 step."""
    yb, acc, xs, level = saved
    T, B, D = grad_y_path.shape
    dev = grad_y_path.device
    eng.backward_begin(B, stages_per_flush=4)
    grad_rows = grad_y_path if (grad_y_path.is_contiguous() and grad_y_path.dtype == torch.float32) else grad_y_path.contiguous().float()
    lam = blocked_zeros(B, D, dev)          # dL/dy_{n+1} through the later steps (without the row's own gradient)
    G_y0 = blocked_zeros(B, D, dev)
    G_a = [blocked_zeros(B, eng.P, dev) for _ in range(4)]
    gx = [blocked_zeros(B, D, dev) for _ in range(4)]
    for n in range(T - 2, -1, -1):
        t0, dt = float(t_host[n]), float(t_host[n + 1]) - float(t_host[n])
        cb = RK38.combo(RK38.b, dt)
        sources = [(grad_rows[n + 1], cb)] + ([(lam, cb)] if n < T - 2 else [])
        eng.combine_backward_multi(sources, B, G_y0, G_a, accumulate=False)
        times = [t0, t0 + RK38.c[1] * dt, t0 + RK38.c[2] * dt, float(t_host[n + 1])]
        stages_backward(eng, RK38, B, yb[n], [acc[n][j] for j in range(3)], times, dt, G_a, gx, 0, 3, xs[n] if level else None, level,
                        y0_accum=G_y0, x_first=0)
        eng.flush()
        lam, G_y0 = G_y0, lam
    rows_block(grad_rows[0], lam, accumulate=True)
    gw = eng.backward_end()
    return rows_unblock(lam, B, D), gw


# --------------------------------------------------------------------------------------------------------
# adaptive dopri5 (torchdiffeq dopri5.py + rk_common.py RKAdaptiveStepsizeODESolver) with the discrete adjoint
# --------------------------------------------------------------------------------------------------------
@dataclass
class _Dopri5Step:
    yb: torch.Tensor                 # blocked state at the start of the accepted step
    A: List[torch.Tensor]            # blocked a_1..a_7 of the step (a_7 = FSAL evaluation at the step's end)
    t0: float
    dt: float
    outputs: List                    # [(row index k into y_path, x = (t_k - t0) / dt)]
    x: Optional[torch.Tensor] = None # what the forward attempt saved for the backward pass of stages 2..7 (operand images), or None
    save_level: int = 0              # 1: stage inputs; 2: + hidden activations and ReLU masks


class Dopri5Stats:
    def __init__(self):
        self.n_accepted = self.n_rejected = self.n_evals = 0


def _cast_time(v: float, time_dtype) -> float:
    if time_dtype == torch.float32:
        return float(np.float32(v))
    return float(v)


def dopri5_forward(eng: TcEngine, y0: torch.Tensor, t_host: Sequence[float], rtol: float, atol: float, *, first_step=None,
                   safety: float = 0.9, ifactor: float = 10.0, dfactor: float = 0.2, max_num_steps: int = 2 ** 31 - 1,
                   time_dtype=torch.float64, save_steps: bool = False, stats: Optional[Dopri5Stats] = None, fp16_forward=None,
                   forward_operands="fp16x2", error_norm: str = "shard", group=None, saved_operands: str = "all"):
    """y0 row-major [B, D] -> y_path [T, B, D] (dense output at the requested times), and the accepted steps when
    `save_steps`.  One host read of the squared-error sum per attempted step decides accept / reject.
    error_norm="global": when agents are sharded over ranks, the squared-error sum and the element count are all-reduced
    (2 doubles per attempt) so that every rank takes the step sequence a single process would take on the whole batch --
    torchdiffeq's RMS norm runs over ALL agents; "shard" uses the local agents only.
    saved_operands (training forward in the split-activation format): what an accepted attempt keeps for the backward pass
    besides (y, a_j): "all" = stage inputs, hidden activations and ReLU masks as the backward kernels' operand images (1,712 B per
    agent-evaluation; the backward kernel recomputes nothing and (y, a_j) of the step are NOT kept), "inputs" = stage inputs only
    (352 B), "none" = nothing (the backward pass rebuilds and recomputes from (y, a_j))."""
    import torch.distributed as dist
    use_global = error_norm == "global" and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    n_elems = torch.tensor([float(y0.shape[0] * y0.shape[1])], dtype=torch.float64, device=y0.device)
    if use_global:
        dist.all_reduce(n_elems, group=group)
    n_elems = float(n_elems.item())
    B, T = y0.shape[0], len(t_host)
    dev = y0.device
    D, P = eng.D, eng.P
    stats = stats if stats is not None else Dopri5Stats()
    # forward_operands: "fp16x2" (default; split activations: the step sequence of the fp32 solver), "fp16", "bf16";
    # the legacy boolean `fp16_forward` (True -> "fp16", False -> "bf16") still wins when given
    prev_fmt = eng.fwd_format
    eng.fwd_format = fwd_format_code(forward_operands if fp16_forward is None else bool(fp16_forward))
    ts = [_cast_time(float(v), time_dtype) for v in t_host]
    y_path = torch.empty((T, B, D), dtype=torch.float32, device=dev)
    y_path[0].copy_(y0)
    y_cur = rows_block(y0)
    y_next = blocked_zeros(B, D, dev)
    A = [blocked_zeros(B, P, dev) for _ in range(7)]
    out_b = blocked_zeros(B, D, dev)
    sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
    zero = Combo(0.0, [], [])
    eng.stage_forward(y_cur, [], zero, ts[0], B, a_out=A[0])
    stats.n_evals += 1

    # ---- initial step size (torchdiffeq misc.py _select_initial_step; Hairer, Norsett & Wanner II.4)
    if first_step is None:
        def rms(x):
            v = torch.stack([x.double().pow(2).sum(), torch.tensor(float(x.numel()), dtype=torch.float64, device=x.device)])
            if use_global:
                dist.all_reduce(v, group=group)
            return float((v[0] / v[1]).sqrt())
        y0f = y0.float()
        f0 = torch.cat([y0f[:, P:2 * P], rows_unblock(A[0], B, P), torch.zeros(B, D - 2 * P, device=dev)], dim=1)
        scale = atol + y0f.abs() * rtol
        d0, d1 = rms(y0f / scale), rms(f0 / scale)
        h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        eng.stage_forward(y_cur, [A[0]], Combo(h0, [0.0], [h0]), _cast_time(ts[0] + h0, time_dtype), B, a_out=A[1])
        stats.n_evals += 1
        f1 = torch.cat([y0f[:, P:2 * P] + h0 * f0[:, P:2 * P], rows_unblock(A[1], B, P), torch.zeros(B, D - 2 * P, device=dev)], dim=1)
        d2 = rms((f1 - f0) / scale) / h0
        h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / 5.0)
        dt = _cast_time(min(100 * h0, h1), time_dtype)
    else:
        dt = _cast_time(float(first_step), time_dtype)

    steps: List[_Dopri5Step] = []
    t0 = t1 = ts[0]
    k = 1
    n_steps = 0
    # The error norm is read on the host after every attempt; whatever the host does between that read and the next launch is
    # GPU idle time.  So: the attempt's ~300 tableau coefficients are assembled in C (ab200_dopri5_attempt), the squared-error
    # sums come from a pre-zeroed pool (no memset launch per attempt), the buffers of the next accepted step are one allocation,
    # and the dense-output rows of an accepted step are launched AFTER the next attempt (their host work overlaps its execution).
    pool_n = 256
    pool = torch.zeros(pool_n, dtype=torch.float64, device=dev)
    pending = None                              # (y, A, dt, outs) of the last accepted step whose dense rows are not launched yet
    side = done = host_sumsq = None
    if dev.type == "cuda":
        side, done = _side_stream(dev), torch.cuda.Event()
        host_sumsq = torch.empty(1, dtype=torch.float64, pin_memory=True)

    def flush_rows():
        nonlocal pending
        if pending is not None:
            yb_, A_, dt_, outs_ = pending
            eng.dopri5_dense_rows(yb_, A_, dt_, [x for _, x in outs_], B, [y_path[kk] for kk, _ in outs_])
            pending = None

    Bp = padded_rows(B)
    # A training forward in the split-activation format also keeps every stage INPUT of the accepted steps as the bf16 operand
    # image of the backward kernels (352 B per agent-stage): the backward pass is bound by its HBM traffic, and loading that
    # image replaces re-reading y0 and up to six a_j (1.5 KB) and spilling it for the weight-gradient kernel.
    if saved_operands not in SAVE_LEVELS:
        raise ValueError("saved_operands must be one of %s, got %r" % (sorted(SAVE_LEVELS), saved_operands))
    save_level = SAVE_LEVELS[saved_operands] if (save_steps and eng.fwd_format == 2) else 0
    xb_floats = eng.xblob_bytes(B, save_level) * 6 // 4 if save_level else 0
    x_cur = torch.empty(xb_floats, dtype=torch.float32, device=dev) if xb_floats else None
    # level 2: the backward pass never reads (y, a_j) of a step (only y of the very first one, which lives outside the sets), so
    # the attempts rotate through THREE buffer sets: the dense-output rows of step n are launched after attempt n + 1 and read
    # y / a_1 = a_7 of step n - 1 and a_2..a_7 of step n; set (n mod 3) is next written by attempt n + 3.
    sets, ci = None, 0
    if save_level == 2:
        sets = [(y_next, A[1:])] + [(blocked_zeros(B, D, dev), [blocked_zeros(B, P, dev) for _ in range(6)]) for _ in range(2)]
    while k < T:
        # ---- one attempted step from t1 with size dt
        assert n_steps < max_num_steps, "max_num_steps exceeded ({}>={})".format(n_steps, max_num_steps)
        assert _cast_time(t1 + dt, time_dtype) > t1, "underflow in dt {}".format(dt)
        ta = t1
        tb = _cast_time(ta + dt, time_dtype)
        slot = n_steps % pool_n
        if slot == 0 and n_steps > 0:
            pool.zero_()
        sumsq = pool[slot:slot + 1]
        eng.dopri5_attempt(y_cur, A, ta, dt, B, y_next, sumsq, rtol, atol, x_cur, save_level)
        stats.n_evals += 6
        if side is not None:
            done.record()            # right behind the attempt, BEFORE the dense rows of the previous step
        flush_rows()
        if side is not None:
            # The error norm travels on a side stream that waits for the attempt only: the host decides accept / reject and
            # launches the next attempt while the dense-output rows of the previous step (~0.2 ms) still run on the main stream,
            # so its turn-around (~50 us, more with the all-reduce of the global norm) is not GPU idle time.
            with torch.cuda.stream(side):
                side.wait_event(done)
                if use_global:
                    dist.all_reduce(sumsq, group=group)
                host_sumsq.copy_(sumsq, non_blocking=True)
            side.synchronize()
            ssq = float(host_sumsq[0])
        else:
            if use_global:
                dist.all_reduce(sumsq, group=group)
            ssq = float(sumsq.item())
        ratio = math.sqrt(ssq / n_elems)      # ONE device->host read per attempt (NaN stays NaN)
        if ratio != ratio:
            eng.check_status()       # a kernel whose bounded barrier wait expired poisons the norm: report that, not an overflow
            raise _lib.Ab200Error("dopri5: non-finite error estimate (state or drift overflowed)")
        n_steps += 1
        if ratio <= 1.0:
            stats.n_accepted += 1
            outs = []
            while k < T and ts[k] <= tb:
                outs.append((k, (ts[k] - ta) / (tb - ta)))
                k += 1
            if save_steps:
                if outs:
                    pending = (y_cur, A, dt, outs)
                steps.append(_Dopri5Step(y_cur, A, ta, dt, outs, x_cur, save_level))
                if save_level == 2:
                    ci = (ci + 1) % 3
                    y_cur, y_next = y_next, sets[ci][0]
                    A = [A[6]] + sets[ci][1]
                    x_cur = torch.empty(xb_floats, dtype=torch.float32, device=dev)
                else:
                    n_state = Bp * (D + 6 * P)
                    buf = torch.empty(n_state + xb_floats, dtype=torch.float32, device=dev)  # y_next + a_2..a_7 (+ X blobs) of the next step
                    y_cur, y_next = y_next, buf[:Bp * D]
                    A = [A[6]] + [buf[Bp * (D + i * P):Bp * (D + (i + 1) * P)] for i in range(6)]
                    x_cur = buf[n_state:] if xb_floats else None
            else:
                if outs:       # buffers are recycled by the next attempt: every requested time inside (t, t + dt] now
                    eng.dopri5_dense_rows(y_cur, A, dt, [x for _, x in outs], B, [y_path[kk] for kk, _ in outs])
                y_cur, y_next = y_next, y_cur
                A[0], A[6] = A[6], A[0]
            t0, t1 = ta, tb
        else:
            stats.n_rejected += 1
        # ---- next step size (misc.py _optimal_step_size)
        if ratio == 0.0:
            dt = _cast_time(dt * ifactor, time_dtype)
        else:
            dfac = 1.0 if ratio < 1.0 else dfactor
            dt = _cast_time(dt * min(ifactor, max(safety / ratio ** 0.2, dfac)), time_dtype)
    flush_rows()
    eng.fwd_format = prev_fmt
    eng.check_status()
    return y_path, (steps if save_steps else None), stats


def dopri5_backward(eng: TcEngine, steps: List[_Dopri5Step], grad_y_path: torch.Tensor):
    """Discrete adjoint of the accepted steps (step sizes are constants, as in torchdiffeq where the controller runs
    under no_grad) -> (grad_y0 row-major, grad_w_flat).  The FSAL evaluation k_7 of a step is the k_1 of the next one
    (one forward evaluation): the next step hands the gradient w.r.t. its k_1 back (`lam_a`) and this step
    differentiates the evaluation once, as its stage 7; only the very first step differentiates its own stage 1."""
    T, B, D = grad_y_path.shape
    dev = grad_y_path.device
    P = eng.P
    eng.backward_begin(B, stages_per_flush=7)
    lam = blocked_zeros(B, D, dev)          # dL/dy at the end of the step being processed
    lam_a = None                            # dL/da_7 handed over by the following step (None: nothing depends on it)
    lam_a_buf = [blocked_zeros(B, P, dev) for _ in range(2)]
    G_y0 = blocked_zeros(B, D, dev)
    G_a = [blocked_zeros(B, P, dev) for _ in range(7)]
    gx = [blocked_zeros(B, D, dev) for _ in range(7)]
    grad_rows = grad_y_path if (grad_y_path.is_contiguous() and grad_y_path.dtype == torch.float32) else grad_y_path.contiguous().float()
    for si in range(len(steps) - 1, -1, -1):
        st = steps[si]
        dt = st.dt
        # step-level gradients: the end state y1 = y0 + dt sum c_sol k, and every dense-output row inside the step (the rows of
        # dL/dy_path are read in place, row-major)
        sources = [(lam, DOPRI5.combo(DOPRI5.b, dt))]
        for (k, x) in st.outputs:
            sources.append((grad_rows[k], DOPRI5.combo(dopri5_interp_weights(x), dt)))
        eng.combine_backward_multi(sources, B, G_y0, G_a, accumulate=False, add_a=lam_a, add_index=6)
        last = 6 if (st.outputs or lam_a is not None) else 5      # stage 7 only matters if something used k_7
        first = 0 if si == 0 else 1                                # k_1 of a later step belongs to the previous step
        combos = [DOPRI5.stage_input(i, dt) for i in range(7)]
        times = [st.t0 + DOPRI5.c[i] * dt for i in range(7)]
        # the stage kernel adds every stage's contribution to dL/dy0 into G_y0 itself and, as one more entry of the same launch,
        # assembles the gradient handed to the previous step's FSAL evaluation: no gather pass over the six gx buffers
        lam_a_next = lam_a_buf[si % 2] if first == 1 else None
        stages_backward(eng, DOPRI5, B, st.yb, st.A, times, dt, G_a, gx, first, last, st.x, st.save_level, y0_accum=G_y0,
                        fsal_out=lam_a_next)
        lam_a = lam_a_next
        eng.flush()
        lam, G_y0 = G_y0, lam
    rows_block(grad_y_path[0], lam, accumulate=True)
    gw = eng.backward_end()
    return rows_unblock(lam, B, D), gw


def _step_backward_n(eng, tab, B, yn, A, stage_times, dt, G_y0_base, G_a_base, gx, out, n_stage):
    combos = [tab.stage_input(i, dt) for i in range(n_stage)]
    for i in range(n_stage - 1, -1, -1):
        later = [l for l in range(i + 1, n_stage) if combos[l].cpa[i] != 0.0 or combos[l].cva[i] != 0.0]
        eng.stage_backward(yn, [A[j] for j in range(i)], combos[i], stage_times[i], B, G_a_base[i], [gx[l] for l in later],
                           [combos[l].cpa[i] for l in later], [combos[l].cva[i] for l in later], gx[i])
    eng.adjoint_gather(G_y0_base, [gx[i] for i in range(n_stage)], [combos[i].cpv for i in range(n_stage)], B, out)
