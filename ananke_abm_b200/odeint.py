"""`odeint` / `odeint_adjoint` with torchdiffeq's call signature -- the solver seam of the reference.

The reference calls (all under /root/reference/src/ananke_abm/models/):
    odeint(self.odefunc, y0, times_union, method="rk4", rtol=..., atol=...)        mode_sep/architecture/model.py:184-191
    odeint(self.ode_func, y0, times, method='dopri5', options={'dtype': float32})  latent_ode/architecture/model.py:192,196
    odeint_adjoint(wrapped_func, x0, t, rtol=, atol=, method='dopri5')             latent_ode/architecture/ode_components.py:50
resolve to this module when it is installed as `torchdiffeq`; the unmodified reference
model files then run on the CUDA kernels.

Dispatch:
  * `func` is one of the two reference drift modules (recognised structurally, see `drift.describe_drift`)
      rk4    -> one fused whole-trajectory launch  (ab200_rk4_forward / ab200_rk4_backward)
      dopri5 -> drift evaluated by `ab200_drift_eval`, stages combined by `ab200_rk_combine_errnorm`
  * any other `func` -> func evaluated by the caller's own torch code on the GPU, every stage combine and the
    error norm by the fused elementwise kernels (`ab200_rk_stage_combine`, `ab200_rk_combine_errnorm`).
Everything runs on CUDA tensors; there is no CPU path (CPU tensors raise).
"""
from __future__ import annotations

import ctypes as C
import warnings
import weakref
from typing import Callable, Optional, Sequence

import torch

from . import _lib
from .drift import DriftSpec, describe_drift

_DEFAULT_PRECISION = {"value": "f32"}


def set_default_precision(p: str) -> None:
    """'f32' (strict FFMA, 1e-5 parity) or 'bf16' (tcgen05 tensor-core path, stated tolerance)."""
    if p not in _lib.PRECISIONS:
        raise ValueError(f"unknown precision {p!r}")
    _DEFAULT_PRECISION["value"] = p


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise _lib.Ab200Error(f"{what} must be a CUDA tensor: ananke_abm_b200 has no CPU path")


_T_CACHE: dict = {}


def _check_t(t: torch.Tensor) -> torch.Tensor:
    """torchdiffeq's input checks (odeint.py `_check_inputs`).  The host copy of a time grid is cached per
    live tensor object and version (a data_ptr key can alias a freed tensor's memory) so that repeated solves on the same grid do not force a device->host sync each call."""
    assert isinstance(t, torch.Tensor), "t must be a torch.Tensor"
    assert t.ndimension() == 1, "t must be one dimensional"
    assert torch.is_floating_point(t), "t must be a floating point Tensor"
    key = id(t)
    hit = _T_CACHE.get(key)
    if hit is not None and hit[0]() is t and hit[1] == t._version:       # same live tensor object, not modified since
        return hit[2]
    t_host = t.detach().to("cpu", non_blocking=False)
    if t_host.numel() > 1:
        d = t_host[1:] - t_host[:-1]
        if not (bool((d > 0).all()) or bool((d < 0).all())):
            raise AssertionError("t must be strictly increasing or decreasing")
    if len(_T_CACHE) > 64:
        _T_CACHE.clear()
    _T_CACHE[key] = (weakref.ref(t), t._version, t_host)
    return t_host


def rk4_forward_into(spec: DriftSpec, w_flat: torch.Tensor, y0: torch.Tensor, t: torch.Tensor, y_path: torch.Tensor,
                     workspace: torch.Tensor, precision: int) -> None:
    """Raw fused launch into caller-owned buffers (no allocation, no sync, no autograd): the call the bench times."""
    L = _lib.lib()
    B, D = y0.shape
    rc = L.ab200_rk4_forward(C.byref(spec.desc), w_flat.data_ptr(), y0.data_ptr(), t.data_ptr(), None, B, t.numel(),
                             y_path.data_ptr(), workspace.data_ptr(), workspace.numel(), precision, _stream_ptr())
    _lib.check(rc, "ab200_rk4_forward")


def rk4_workspace(spec: DriftSpec, B: int, T: int, precision: int, device) -> torch.Tensor:
    nbytes = _lib.lib().ab200_rk4_workspace_bytes(C.byref(spec.desc), B, T, precision)
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------------------------
# fused rk4 for the recognised drift nets
# --------------------------------------------------------------------------------------------------
class _FusedRK4(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, t, w_flat, spec: DriftSpec, precision: int, t_host):
        L = _lib.lib()
        B, D = y0.shape
        T = t.numel()
        y0c = y0.contiguous().float()
        tc = t.contiguous().float()
        wc = w_flat.contiguous().float()
        th = t_host.contiguous().float()
        n_par = L.ab200_drift_param_count(C.byref(spec.desc))
        if wc.numel() != n_par:
            raise _lib.Ab200Error(f"drift parameter vector has {wc.numel()} elements, the descriptor needs {n_par}")
        y_path = torch.empty((T, B, D), dtype=torch.float32, device=y0.device)
        nbytes = L.ab200_rk4_workspace_bytes(C.byref(spec.desc), B, T, precision)
        ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=y0.device)
        rc = L.ab200_rk4_forward(C.byref(spec.desc), wc.data_ptr(), y0c.data_ptr(), tc.data_ptr(), th.data_ptr(), B, T,
                                 y_path.data_ptr(), ws.data_ptr(), ws.numel(), precision, _stream_ptr())
        _lib.check(rc, "ab200_rk4_forward")
        if precision == _lib.PREC_BF16:
            # the tensor-core trajectory kernel keeps a status word in the last 256 bytes of its workspace: a bounded barrier
            # wait that expired means the trajectory is garbage (one 4-byte host read per solve)
            if int(ws[ws.numel() - 256:ws.numel() - 252].view(torch.int32).item()) != 0:
                raise _lib.Ab200Error("rk4 tensor-core kernel: barrier wait timed out, results discarded")
        ctx.spec, ctx.precision = spec, precision
        ctx.save_for_backward(tc, wc, y_path)
        return y_path

    @staticmethod
    def backward(ctx, grad_y_path):
        L = _lib.lib()
        tc, wc, y_path = ctx.saved_tensors
        spec = ctx.spec
        T, B, D = y_path.shape
        g = grad_y_path.contiguous().float()
        gy0 = torch.empty((B, D), dtype=torch.float32, device=g.device)
        gw = torch.empty_like(wc)
        # gradients are always taken on the strict-fp32 path unless a tensor-core backward exists for `precision`
        prec = ctx.precision if ctx.precision in spec.backward_precisions else _lib.PREC_F32
        nbytes = L.ab200_rk4_backward_workspace_bytes(C.byref(spec.desc), B, T, prec)
        ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=g.device)
        rc = L.ab200_rk4_backward(C.byref(spec.desc), wc.data_ptr(), tc.data_ptr(), y_path.data_ptr(), g.data_ptr(), B, T,
                                  gy0.data_ptr(), gw.data_ptr(), ws.data_ptr(), ws.numel(), prec, _stream_ptr())
        _lib.check(rc, "ab200_rk4_backward")
        return gy0, None, gw, None, None, None


class _StageRK4TC(torch.autograd.Function):
    """rk4 on the tensor-core STAGE kernels with the discrete adjoint (bf16 operands, fp32 accumulation and state):
    the training path of `precision='bf16'`.  Saves the trajectory rows and the three stage accelerations per step."""

    @staticmethod
    def forward(ctx, y0, t, w_flat, spec: DriftSpec, t_host, saved_operands="none"):
        from . import stage
        eng = stage.TcEngine(spec, w_flat)
        th = [float(v) for v in t_host.tolist()]
        y_path, (yb, acc, xs, level) = stage.rk4_forward(eng, y0.contiguous().float(), th, save_stages=True, saved_operands=saved_operands)
        ctx.eng, ctx.th, ctx.level = eng, th, level
        ctx.save_for_backward(yb, acc, *([xs] if xs is not None else []))
        return y_path

    @staticmethod
    def backward(ctx, grad_y_path):
        from . import stage
        yb, acc, *rest = ctx.saved_tensors
        gy0, gw = stage.rk4_backward(ctx.eng, ctx.th, (yb, acc, rest[0] if rest else None, ctx.level), grad_y_path.contiguous().float())
        return gy0, None, gw, None, None, None


class _StageDopri5TC(torch.autograd.Function):
    """dopri5 on the tensor-core STAGE kernels: adaptive forward with dense output, discrete adjoint of the accepted
    steps in backward (what autograd through torchdiffeq's solver ops computes; the step-size controller is constant,
    as it runs under no_grad there)."""

    @staticmethod
    def forward(ctx, y0, t, w_flat, spec: DriftSpec, t_host, rtol, atol, opts):
        from . import stage
        eng = stage.TcEngine(spec, w_flat)
        th = [float(v) for v in t_host.tolist()]
        need = torch.is_grad_enabled() or y0.requires_grad or w_flat.requires_grad
        stats = stage.Dopri5Stats()
        y_path, steps, _ = stage.dopri5_forward(eng, y0.contiguous().float(), th, rtol, atol, save_steps=need, stats=stats, **opts)
        ctx.eng, ctx.steps = eng, steps
        _LAST["solver"] = stats
        return y_path

    @staticmethod
    def backward(ctx, grad_y_path):
        from . import stage
        if ctx.steps is None:
            raise RuntimeError("dopri5 (tensor-core stage path): the saved steps were freed by the first backward pass; "
                               "a second backward through the same solve is not supported (re-run the forward)")
        gy0, gw = stage.dopri5_backward(ctx.eng, ctx.steps, grad_y_path.contiguous().float())
        ctx.steps = None
        return gy0, None, gw, None, None, None, None, None


def drift_eval(spec: DriftSpec, w_flat: torch.Tensor, t: float, y: torch.Tensor, precision: int = 0) -> torch.Tensor:
    """One evaluation f(t, y) of a recognised drift net on the CUDA path (no autograd)."""
    L = _lib.lib()
    _require_cuda(y, "y")
    B = y.shape[0]
    n_par = L.ab200_drift_param_count(C.byref(spec.desc))
    if w_flat.numel() != n_par:
        raise _lib.Ab200Error(f"drift parameter vector has {w_flat.numel()} elements, the descriptor needs {n_par}")
    out = torch.empty_like(y, dtype=torch.float32)
    nbytes = L.ab200_drift_eval_workspace_bytes(C.byref(spec.desc), B, precision)
    ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=y.device)
    rc = L.ab200_drift_eval(C.byref(spec.desc), w_flat.contiguous().float().data_ptr(), float(t), y.contiguous().float().data_ptr(),
                            B, out.data_ptr(), ws.data_ptr(), ws.numel(), precision, _stream_ptr())
    _lib.check(rc, "ab200_drift_eval")
    return out


def drift_vjp(spec: DriftSpec, w_flat: torch.Tensor, t: float, y: torch.Tensor, grad_out: torch.Tensor):
    """J^T grad_out of one drift evaluation at (t, y) in strict fp32 (ab200_drift_vjp) -> (grad_y [B, D], grad_w_flat)."""
    L = _lib.lib()
    _require_cuda(y, "y")
    B = y.shape[0]
    yc, gc, wc = y.contiguous().float(), grad_out.contiguous().float(), w_flat.contiguous().float()
    gy = torch.empty_like(yc)
    gw = torch.empty_like(wc)
    nbytes = L.ab200_drift_vjp_workspace_bytes(C.byref(spec.desc), B)
    ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=y.device)
    rc = L.ab200_drift_vjp(C.byref(spec.desc), wc.data_ptr(), float(t), yc.data_ptr(), gc.data_ptr(), B, gy.data_ptr(), gw.data_ptr(),
                           ws.data_ptr(), ws.numel(), _stream_ptr())
    _lib.check(rc, "ab200_drift_vjp")
    return gy, gw


class _DriftFn(torch.autograd.Function):
    """f(t, y) of a recognised drift on the strict-fp32 kernels WITH a backward: `ab200_drift_eval` forward,
    `ab200_drift_vjp` backward.  This is what makes a solver written in PyTorch ops over the kernel-evaluated drift
    differentiable (autograd through dopri5, the reference's training path) and what the continuous adjoint evaluates."""

    @staticmethod
    def forward(ctx, y, w_flat, spec: DriftSpec, t: float):
        yd, wd = y.detach(), w_flat.detach()
        ctx.spec, ctx.t = spec, float(t)
        ctx.save_for_backward(yd, wd)
        return drift_eval(spec, wd, float(t), yd)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        y, w = ctx.saved_tensors
        gy, gw = drift_vjp(ctx.spec, w, ctx.t, y, g)
        return gy, gw, None, None


def drift_apply(spec: DriftSpec, t, y: torch.Tensor) -> torch.Tensor:
    """One evaluation of a recognised drift through the kernels; differentiable (first order) w.r.t. y and the module's
    parameters whenever autograd is recording."""
    tf = float(t)
    if torch.is_grad_enabled() and (y.requires_grad or any(p.requires_grad for p in spec.params)):
        return _DriftFn.apply(y, spec.flat_params(), spec, tf)
    return drift_eval(spec, spec.flat_params().detach(), tf, y)


# --------------------------------------------------------------------------------------------------
# fused elementwise stage algebra for arbitrary `func`
# --------------------------------------------------------------------------------------------------
class _StageCombine(torch.autograd.Function):
    """out = y + dt * sum_j coef[j] * k_j   in one pass (ab200_rk_stage_combine)."""

    @staticmethod
    def forward(ctx, dt: float, coef: Sequence[float], y, *ks):
        L = _lib.lib()
        yc = y.contiguous()
        kc = [k.contiguous() for k in ks]
        out = torch.empty_like(yc)
        n_k = len(kc)
        ptrs = (C.c_void_p * max(n_k, 1))(*[k.data_ptr() for k in kc])
        cf = (C.c_float * max(n_k, 1))(*[float(c) for c in coef])
        rc = L.ab200_rk_stage_combine(yc.data_ptr(), C.cast(ptrs, C.c_void_p), C.cast(cf, C.c_void_p), n_k, float(dt),
                                      out.data_ptr(), yc.numel(), _stream_ptr())
        _lib.check(rc, "ab200_rk_stage_combine")
        ctx.dt, ctx.coef = float(dt), [float(c) for c in coef]
        return out

    @staticmethod
    def backward(ctx, g):
        return (None, None, g) + tuple(g * (ctx.dt * c) for c in ctx.coef)


def _combine(y, ks, coef, dt):
    return _StageCombine.apply(float(dt), list(coef), y, *ks)


def _rk4_generic(func, y0, t, t_host, time_as_float: bool = False):
    third, two_thirds = 1.0 / 3.0, 2.0 / 3.0
    sol = [y0]
    y = y0
    if time_as_float:       # kernel-evaluated func: stage times as host scalars, computed in the state's dtype like the tensor path
        t = t_host.to(y0.dtype)
    for i in range(t.numel() - 1):
        t0, t1 = t[i], t[i + 1]
        dt_t = t1 - t0
        dt = float(t_host[i + 1] - t_host[i])
        k1 = func(t0, y)
        k2 = func(t0 + dt_t * third, _combine(y, [k1], [third], dt))
        k3 = func(t0 + dt_t * two_thirds, _combine(y, [k1, k2], [-third, 1.0], dt))
        k4 = func(t1, _combine(y, [k1, k2, k3], [1.0, -1.0, 1.0], dt))
        y = _combine(y, [k1, k2, k3, k4], [0.125, 0.375, 0.375, 0.125], dt)
        sol.append(y)
    return torch.stack(sol, dim=0)


# Dormand-Prince 5(4), Shampine's error weights -- torchdiffeq dopri5.py
_DP_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_DP_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_DP_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
_DP_C_ERR = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
             -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0]
_DP_C_MID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
             187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]


class _Dopri5:
    """Adaptive Dormand-Prince driver (tdq rk_common.py RKAdaptiveStepsizeODESolver).  The drift is evaluated by
    `f`; stage combines and the RMS error norm are single fused passes over the state.  The accept/reject
    decision reads ONE float per attempted step from the device (the squared-error sum)."""

    def __init__(self, f: Callable, y0, rtol, atol, first_step=None, safety=0.9, ifactor=10.0, dfactor=0.2,
                 max_num_steps=2 ** 31 - 1, dtype=torch.float64, norm=None, time_as_float: bool = False, segments=None, **unused):
        if unused:
            warnings.warn(f"Dopri5: Unexpected arguments {unused}")
        if norm is not None:
            raise NotImplementedError("custom norms are not supported on the fused path")
        self.f, self.y0 = f, y0
        self.rtol, self.atol = float(rtol), float(atol)
        self.safety, self.ifactor, self.dfactor = float(safety), float(ifactor), float(dfactor)
        self.first_step = first_step
        self.max_num_steps = int(max_num_steps)
        self.tdtype = torch.promote_types(dtype, y0.dtype)
        self.n_accepted = self.n_rejected = 0
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=y0.device)
        self.time_as_float = bool(time_as_float)      # kernel-evaluated drifts take the time as a host scalar (no device round trip)
        # torchdiffeq's mixed norm (misc.py `_mixed_norm`: max over the RMS norms of the tuple components) for a flat state that
        # packs several components: `segments` = [(offset, length), ...]; None = one RMS norm over the whole state
        self.segments = None if segments is None else [(int(o), int(n)) for o, n in segments if int(n) > 0]
        if self.segments is not None:
            self._sumsq = torch.zeros(len(self.segments), dtype=torch.float32, device=y0.device)

    def _tt(self, v: float):
        if self.time_as_float:
            return float(torch.tensor(v, dtype=self.y0.dtype)) if self.y0.dtype != torch.float64 else float(v)
        return torch.tensor(v, dtype=self.y0.dtype, device=self.y0.device)

    def _cast(self, v: float) -> float:
        """round a time-like python float to the solver's time dtype (float32 when options['dtype']=float32)."""
        if self.tdtype == torch.float32:
            return float(torch.tensor(v, dtype=torch.float32))
        return float(v)

    def _rms(self, x: torch.Tensor) -> float:
        x = x.detach()
        if self.segments is None:
            return float(x.float().pow(2).mean().sqrt())
        flat = x.reshape(-1).float()
        return float(torch.stack([flat[o:o + n].pow(2).mean() for o, n in self.segments]).max().sqrt())

    def _initial_step(self, t0: float, f0):
        y0 = self.y0
        scale = self.atol + y0.abs() * self.rtol
        d0, d1 = self._rms(y0 / scale), self._rms(f0 / scale)
        h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        y1 = _combine(y0, [f0], [1.0], h0)
        f1 = self.f(self._tt(self._cast(t0 + h0)), y1)
        d2 = self._rms((f1 - f0) / scale) / h0
        if d1 <= 1e-15 and d2 <= 1e-15:
            h1 = max(1e-6, h0 * 1e-3)
        else:
            h1 = (0.01 / max(d1, d2)) ** (1.0 / 5.0)
        return self._cast(min(100 * h0, h1))

    def _step(self, y0, f0, t0: float, dt: float):
        L = _lib.lib()
        ks = [f0]
        yi = y0
        t1 = self._cast(t0 + dt)
        for al, be in zip(_DP_ALPHA, _DP_BETA):
            ti = t1 if al == 1.0 else self._cast(t0 + al * dt)
            yi = _combine(y0, ks, be, dt)
            ks.append(self.f(self._tt(ti), yi))
        y1, f1 = yi, ks[-1]                     # FSAL: the last stage input is the 5th-order solution
        # embedded error estimate and its RMS norm in one pass over the state (y1_out = NULL: not re-stored)
        n = y0.numel()
        kc = [k.contiguous() for k in ks]
        ptrs = (C.c_void_p * 8)(*([k.data_ptr() for k in kc] + [0]))
        csol = (C.c_float * 8)(*([float(c) for c in _DP_C_SOL] + [0.0]))
        cerr = (C.c_float * 8)(*([float(c) for c in _DP_C_ERR] + [0.0]))
        self._sumsq.zero_()
        y0c = y0.contiguous()
        if self.segments is None:
            rc = L.ab200_rk_combine_errnorm(y0c.data_ptr(), C.cast(ptrs, C.c_void_p), C.cast(csol, C.c_void_p),
                                            C.cast(cerr, C.c_void_p), 7, float(dt), self.rtol, self.atol, None,
                                            self._sumsq.data_ptr(), n, _stream_ptr())
            _lib.check(rc, "ab200_rk_combine_errnorm")
            ratio = float(torch.sqrt(self._sumsq[0] / n))
        else:
            # one error-norm pass per component of the packed state; the ratio is the max of the per-component RMS values
            for i, (o, m) in enumerate(self.segments):
                sp = (C.c_void_p * 8)(*([k.data_ptr() + 4 * o for k in kc] + [0]))
                rc = L.ab200_rk_combine_errnorm(y0c.data_ptr() + 4 * o, C.cast(sp, C.c_void_p), C.cast(csol, C.c_void_p),
                                                C.cast(cerr, C.c_void_p), 7, float(dt), self.rtol, self.atol, None,
                                                self._sumsq.data_ptr() + 4 * i, m, _stream_ptr())
                _lib.check(rc, "ab200_rk_combine_errnorm")
            cnt = torch.tensor([float(m) for _, m in self.segments], dtype=torch.float32, device=y0.device)
            ratio = float(torch.sqrt((self._sumsq / cnt).max()))
        return y1, f1, ratio, ks

    def _next_dt(self, dt: float, ratio: float) -> float:
        if ratio == 0:
            return self._cast(dt * self.ifactor)
        dfac = 1.0 if ratio < 1 else self.dfactor
        factor = min(self.ifactor, max(self.safety / ratio ** 0.2, dfac))
        return self._cast(dt * factor)

    def integrate(self, t_host: torch.Tensor):
        y0 = self.y0
        ts = [self._cast(float(v)) for v in t_host.tolist()]
        f0 = self.f(self._tt(ts[0]), y0)
        dt = self._initial_step(ts[0], f0) if self.first_step is None else self._cast(float(self.first_step))
        t0 = t1 = ts[0]
        y, f = y0, f0
        coeffs = None
        out = [y0]
        for tn in ts[1:]:
            n_steps = 0
            while tn > t1:
                assert n_steps < self.max_num_steps, "max_num_steps exceeded ({}>={})".format(n_steps, self.max_num_steps)
                assert self._cast(t1 + dt) > t1, "underflow in dt {}".format(dt)
                y1, f1, ratio, ks = self._step(y, f, t1, dt)
                if ratio <= 1:
                    self.n_accepted += 1
                    ymid = _combine(y, ks, _DP_C_MID, dt)
                    coeffs = _interp_fit(y, y1, ymid, ks[0], ks[-1], dt)
                    t0, t1 = t1, self._cast(t1 + dt)
                    y, f = y1, f1
                else:
                    self.n_rejected += 1
                dt = self._next_dt(dt, ratio)
                n_steps += 1
            out.append(_interp_eval(coeffs, t0, t1, tn))
        return torch.stack(out, dim=0)


def _interp_fit(y0, y1, y_mid, f0, f1, dt):
    a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
    b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
    c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
    d = dt * f0
    return [y0, d, c, b, a]


def _interp_eval(coeffs, t0, t1, t):
    assert t0 <= t <= t1, "invalid interpolation, fails `t0 <= t <= t1`"
    x = (t - t0) / (t1 - t0)
    total = coeffs[0] + x * coeffs[1]
    xp = x
    for c in coeffs[2:]:
        xp = xp * x
        total = total + xp * c
    return total


_LAST = {"solver": None}


# --------------------------------------------------------------------------------------------------
# public API
# --------------------------------------------------------------------------------------------------
def odeint(func, y0, t, *, rtol: float = 1e-7, atol: float = 1e-9, method: Optional[str] = None,
           options: Optional[dict] = None, event_fn=None):
    """Drop-in for `torchdiffeq.odeint` on the path the reference uses.  Returns `[len(t), *y0.shape]`."""
    if event_fn is not None:
        raise NotImplementedError("event handling is outside the reference's path")
    if isinstance(y0, (tuple, list)):
        raise NotImplementedError("tuple states are outside the reference's path (it passes a single [B, D] tensor)")
    options = {} if options is None else dict(options)
    method = "dopri5" if method is None else method
    _require_cuda(y0, "y0")
    t_host = _check_t(t)
    t = t.to(y0.device)
    if y0.numel() == 0:      # empty batch (a rank with no agents): torchdiffeq returns the stacked, still empty, state
        if options.get("error_norm") == "global":
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(options.get("group")) > 1:
                # the other ranks all-reduce the error norm once per attempt: a rank that returns here would leave them waiting
                raise _lib.Ab200Error("error_norm='global' needs at least one agent on every rank (and the same number of solves "
                                      "per rank): an empty shard cannot take part in the per-attempt all-reduce")
        return y0.unsqueeze(0).repeat(t_host.numel(), *([1] * y0.dim()))
    prec_name = options.pop("precision", _DEFAULT_PRECISION["value"])
    precision = _lib.PRECISIONS[prec_name]

    decreasing = t_host.numel() > 1 and bool(t_host[0] > t_host[1])
    spec = describe_drift(func) if (y0.dim() == 2 and not decreasing) else None

    if method == "rk4":
        step_size = options.pop("step_size", None)
        if step_size is not None and t_host.numel() > 1:
            return _rk4_step_size(func, y0, t, t_host, float(step_size), dict(options, precision=prec_name))
        time_as_float = bool(options.pop("time_as_float", False))
        for k in ("dtype", "norm", "segments", "adjoint_mode"):
            options.pop(k, None)
        # tensor-core training path: what the forward saves for the backward pass (stage.rk4_forward).  Measured per grid step over
        # 250,112 agents (scripts/rk4_levels_time.py): none 0.49 + 2.01 ms, inputs 0.75 + 1.79, all 0.82 + 1.45 -- the single-term fp16
        # forward of "none" is the cheapest, "all" wins 9 % at 5x the saved bytes per agent-step, "inputs" loses: default "none"
        saved_operands = options.pop("saved_operands", "none")
        if options:
            warnings.warn(f"rk4: Unexpected arguments {options}")
        if spec is not None and y0.shape[1] == spec.state_dim and y0.dtype == torch.float32:
            w_flat = spec.flat_params()
            needs_grad = torch.is_grad_enabled() and (y0.requires_grad or w_flat.requires_grad)
            if precision == _lib.PREC_BF16 and needs_grad and spec.tc_stage_supported():
                return _StageRK4TC.apply(y0, t, w_flat, spec, t_host, saved_operands)
            return _FusedRK4.apply(y0, t, w_flat, spec, precision, t_host)
        if y0.dtype != torch.float32:
            raise _lib.Ab200Error(f"rk4: the fused stage-combine kernels are fp32, got y0 of {y0.dtype} (cast the state)")
        f = _wrap_func(func, y0, decreasing)
        tt = -t if decreasing else t
        th = -t_host if decreasing else t_host
        return _rk4_generic(f, y0, tt.to(y0.dtype), th, time_as_float)
    if method == "dopri5":
        # Modules whose forward() is a bare kernel call (this package's mirrors) cannot be differentiated by autograd: a
        # training call on them takes the tensor-core stage path (the only differentiable dopri5 for them) or fails loudly.
        if (spec is not None and y0.shape[1] == spec.state_dim and precision == _lib.PREC_BF16 and spec.tc_stage_supported()
                and y0.dtype == torch.float32):
            opts = {}
            for k_ in ("first_step", "safety", "ifactor", "dfactor", "max_num_steps", "fp16_forward", "forward_operands", "error_norm", "group", "saved_operands"):
                if k_ in options:
                    opts[k_] = options.pop(k_)
            opts["time_dtype"] = torch.promote_types(options.pop("dtype", torch.float64), torch.float32)
            options.pop("norm", None)
            if options:
                warnings.warn(f"dopri5: Unexpected arguments {options}")
            w_flat = spec.flat_params()
            needs_grad = torch.is_grad_enabled() and (y0.requires_grad or w_flat.requires_grad)
            if needs_grad:
                return _StageDopri5TC.apply(y0, t, w_flat, spec, t_host, float(rtol), float(atol), opts)
            with torch.no_grad():
                return _StageDopri5TC.apply(y0, t, w_flat.detach(), spec, t_host, float(rtol), float(atol), opts)
        if spec is not None and y0.shape[1] == spec.state_dim and y0.dtype == torch.float32:
            # strict fp32: every drift evaluation is ab200_drift_eval; when autograd is recording, each evaluation carries
            # ab200_drift_vjp as its backward, so `loss.backward()` through the solver ops (the reference's training path,
            # latent_ode/train/train.py:73) runs on kernels for both reference drift shapes
            f = lambda tt, yy: drift_apply(spec, tt, yy)   # noqa: E731
            options["time_as_float"] = True
        else:
            if y0.dtype != torch.float32:
                raise _lib.Ab200Error(f"dopri5: the fused stage-combine kernels are fp32, got y0 of {y0.dtype} (cast the state)")
            f = _wrap_func(func, y0, decreasing)
        th = -t_host if decreasing else t_host
        solver = _Dopri5(f, y0, rtol, atol, **options)
        _LAST["solver"] = solver
        return solver.integrate(th)
    raise ValueError(f'Invalid method "{method}".')


def _rk4_step_size(func, y0, t, t_host, step_size: float, options: dict):
    """torchdiffeq solvers.py `FixedGridODESolver` with options['step_size']: the solver steps over t[0] + k step_size (last
    point moved onto t[-1]) and every requested time is a LINEAR interpolant of the two grid rows around it (the grid row itself
    where they coincide).  General form: solve on the grid through the path `odeint` would take anyway, then interpolate with
    torch ops (differentiable).  It materialises every grid row; the memory-light variant that keeps two state buffers is the
    tensor-core continuous adjoint (adjoint_tc.py)."""
    import numpy as np
    from .adjoint_tc import step_grid
    th = [float(v) for v in t_host.tolist()]
    npdt = np.float64 if t_host.dtype == torch.float64 else np.float32
    grid = step_grid(th[0], th[-1], step_size, npdt)
    sign = 1.0 if th[-1] >= th[0] else -1.0
    yg = odeint(func, y0, torch.tensor(grid, dtype=t.dtype, device=y0.device), method="rk4", options=options)
    rows, j = [yg[0]], 1
    for n in range(len(grid) - 1):
        t0, t1 = grid[n], grid[n + 1]
        while j < len(th) and sign * t1 >= sign * th[j]:
            if th[j] == t1:
                rows.append(yg[n + 1])
            elif th[j] == t0:
                rows.append(yg[n])
            else:
                rows.append(yg[n] + ((th[j] - t0) / (t1 - t0)) * (yg[n + 1] - yg[n]))
            j += 1
    return torch.stack(rows, dim=0)


def _wrap_func(func, y0, decreasing: bool):
    def conv(tt, yy):      # torchdiffeq casts t to the state's dtype before every user call; host scalars pass through
        return tt.to(yy.dtype) if torch.is_tensor(tt) else tt
    if decreasing:
        return lambda tt, yy: -func(conv(-tt, yy), yy)
    return lambda tt, yy: func(conv(tt, yy), yy)


def odeint_adjoint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None, adjoint_rtol=None,
                   adjoint_atol=None, adjoint_method=None, adjoint_options=None, adjoint_params=None):
    """Drop-in for `torchdiffeq.odeint_adjoint` (the call at latent_ode/architecture/ode_components.py:50).

    Default = torchdiffeq's semantics: forward solve under no_grad (any `options['precision']`), backward by integrating the
    augmented system [y, a_y, a_theta] from t[i] to t[i-1] with adjoint_rtol / adjoint_atol / adjoint_method and the mixed
    error norm, re-seeding y from the saved rows (adjoint.py).  For the two reference drift shapes the augmented dynamics run
    on kernels (`ab200_drift_eval` + `ab200_drift_vjp`, strict fp32).

    Explicit opt-in `options={'adjoint_mode': 'discrete'}`: the discrete adjoint of the accepted steps (what autograd through
    `odeint` computes, i.e. the gradient of the reference's LIVE training paths), on the fused fp32 rk4 adjoint or the tensor-core
    stage path according to `options['precision']`.  It ignores adjoint_* by construction, so passing any of them together
    with this mode is an error rather than a silent override.
    """
    if event_fn is not None:
        raise NotImplementedError("event handling is outside the reference's path")
    if adjoint_params is None and not isinstance(func, torch.nn.Module):
        raise ValueError("func must be an instance of nn.Module to specify the adjoint parameters; alternatively they "
                         "can be specified explicitly via the `adjoint_params` argument.")
    method = "dopri5" if method is None else method
    opt = dict(options or {})
    mode = opt.pop("adjoint_mode", "continuous")
    if mode == "discrete":
        given = [k for k, v in (("adjoint_rtol", adjoint_rtol), ("adjoint_atol", adjoint_atol), ("adjoint_method", adjoint_method),
                                ("adjoint_options", adjoint_options), ("adjoint_params", adjoint_params)) if v is not None]
        if given:
            raise ValueError(f"adjoint_mode='discrete' differentiates the forward steps themselves: {given} would be ignored")
        return odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=opt)
    if mode != "continuous":
        raise ValueError(f"unknown adjoint_mode {mode!r} (choose 'continuous' or 'discrete')")
    from .adjoint import continuous_adjoint
    return continuous_adjoint(func, y0, t, rtol=rtol, atol=atol, method=method, options=opt,
                              adjoint_rtol=adjoint_rtol, adjoint_atol=adjoint_atol, adjoint_method=adjoint_method,
                              adjoint_options=adjoint_options, adjoint_params=adjoint_params)
