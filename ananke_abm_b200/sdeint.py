"""`sdeint` with torchsde's call signature for the reference's SDE branch:

    sdeint(sde, y0, ts, method="euler", dt=0.01, options=...)     latent_ode/architecture/model.py:192-194
                                                                  mode_sep/architecture/model.py:158-182

Fixed-step Euler-Maruyama for diagonal Ito noise: the drift `sde.f` of the recognised modules is evaluated by
`ab200_drift_eval` (any other `sde` by its own torch `f`), the step itself by `ab200_sde_euler_step` (one pass over the
state, counter-based Philox noise generated in the kernel).  Stepping follows torchsde's fixed-step solvers: grid
t0 + k*dt clipped at ts[-1], requested times read off by linear interpolation between the surrounding grid points.

A call that needs gradients records the same steps with autograd (`_EulerMaruyamaStep`: the step kernel forward, d/dy = 1,
d/df = h, d/dg = sqrt(h) xi backward; the drift evaluation carries `ab200_drift_vjp`): the reference's default latent_ode
training path (latent_ode/train/train.py:57-74 with enable_sde=True).  The noise is this package's own specification (csrc/sde_em.cu) -- torchsde's
Brownian interval is not reproducible -- seeded by `seed=` (or drawn from torch's generator when omitted).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .drift import describe_drift


def _drift_module(sde):
    """the module whose forward is the drift: the object itself, or the `.base` of the reference's ScaledSDE wrapper"""
    if describe_drift(sde) is not None:
        return sde
    base = getattr(sde, "base", None)
    if base is not None and describe_drift(base) is not None:
        return base
    return None


class _EulerMaruyamaStep(torch.autograd.Function):
    """y_next = y + f h + g sqrt(h) xi  as ONE kernel pass (ab200_sde_euler_step, xi = counter-based Philox noise of (seed, step))
    with its backward: d/dy = 1, d/df = h, d/dg = sqrt(h) xi.  The drift evaluation feeding `f` carries its own backward
    (ab200_drift_vjp), so autograd through the whole Euler-Maruyama loop -- the reference's default latent_ode training path,
    latent_ode/train/train.py:57-74 with enable_sde=True -- runs on kernels."""

    @staticmethod
    def forward(ctx, y, f, g, h: float, seed: int, step: int):
        L = _lib.lib()
        B, D = y.shape
        yc, fc, gc = y.contiguous().float(), f.contiguous().float(), g.contiguous().float()
        out = torch.empty_like(yc)
        need_xi = ctx.needs_input_grad[2]
        xi = torch.empty_like(yc) if need_xi else None
        rc = L.ab200_sde_euler_step(yc.data_ptr(), fc.data_ptr(), gc.data_ptr(), 1, B, D, float(h), int(seed), int(step), out.data_ptr(),
                                    None if xi is None else xi.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "ab200_sde_euler_step")
        ctx.h = float(h)
        if need_xi:
            ctx.save_for_backward(xi)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        gg = None
        if ctx.needs_input_grad[2]:
            (xi,) = ctx.saved_tensors
            gg = go * (xi * (ctx.h ** 0.5))
        return go, go * ctx.h, gg, None, None, None


def _integrate_with_grad(sde, y0, ts, dt: float, seed: int):
    """the same scheme as `_integrate`, recorded by autograd (training through the sampler)"""
    from .odeint import drift_apply
    t_list = [float(v) for v in ts.tolist()]
    dev = y0.device
    mod = _drift_module(sde)
    spec = describe_drift(mod) if mod is not None else None
    rows = [y0]
    curr_y, prev_y = y0.float(), None
    curr_t = prev_t = t_list[0]
    step = 0
    for out_t in t_list[1:]:
        while curr_t < out_t:
            next_t = min(curr_t + dt, t_list[-1])
            h = float(torch.tensor(next_t - curr_t, dtype=torch.float32))
            tt = torch.tensor(curr_t, dtype=torch.float32, device=dev)
            f = drift_apply(spec, curr_t, curr_y) if spec is not None else sde.f(tt, curr_y).float()
            g = sde.g(tt, curr_y).float()
            if g.shape != curr_y.shape:
                raise ValueError("sdeint: diagonal noise expected (g(t, y) must have the shape of y)")
            prev_t, curr_t = curr_t, next_t
            prev_y, curr_y = curr_y, _EulerMaruyamaStep.apply(curr_y, f, g, h, seed, step)
            step += 1
        if prev_y is None or curr_t == out_t:
            rows.append(curr_y)
        else:
            w = float(torch.tensor((out_t - prev_t) / (curr_t - prev_t), dtype=torch.float32))
            rows.append(torch.lerp(prev_y, curr_y, w))
    return torch.stack(rows, dim=0)


@torch.no_grad()
def _integrate(sde, y0, ts, dt: float, seed: int, return_noise: bool = False):
    from .odeint import drift_eval
    L = _lib.lib()
    B, D = y0.shape
    t_list = [float(v) for v in ts.tolist()]
    dev = y0.device
    mod = _drift_module(sde)
    spec = describe_drift(mod) if mod is not None else None
    w_flat = spec.flat_params().detach() if spec is not None else None
    stream = torch.cuda.current_stream().cuda_stream
    out = torch.empty((len(t_list), B, D), dtype=torch.float32, device=dev)
    out[0] = y0
    curr_y, prev_y = y0.contiguous().float().clone(), None
    nxt = torch.empty_like(curr_y)
    curr_t = prev_t = t_list[0]
    step = 0
    noises = []
    for k, out_t in enumerate(t_list[1:], start=1):
        while curr_t < out_t:
            next_t = min(curr_t + dt, t_list[-1])
            h = float(torch.tensor(next_t - curr_t, dtype=torch.float32))
            tt = torch.tensor(curr_t, dtype=torch.float32, device=dev)
            f = drift_eval(spec, w_flat, curr_t, curr_y) if spec is not None else sde.f(tt, curr_y).contiguous().float()
            g = sde.g(tt, curr_y).contiguous().float()
            if g.shape != curr_y.shape:
                raise ValueError("sdeint: diagonal noise expected (g(t, y) must have the shape of y)")
            xi = torch.empty_like(curr_y) if return_noise else None
            rc = L.ab200_sde_euler_step(curr_y.data_ptr(), f.data_ptr(), g.data_ptr(), 1, B, D, h, seed, step, nxt.data_ptr(),
                                        None if xi is None else xi.data_ptr(), stream)
            _lib.check(rc, "ab200_sde_euler_step")
            if return_noise:
                noises.append(xi)
            prev_t, curr_t = curr_t, next_t
            prev_y, curr_y, nxt = curr_y, nxt, (prev_y if prev_y is not None else torch.empty_like(curr_y))
            step += 1
        if prev_y is None or curr_t == out_t:
            out[k] = curr_y
        else:
            w = float(torch.tensor((out_t - prev_t) / (curr_t - prev_t), dtype=torch.float32))
            torch.lerp(prev_y, curr_y, w, out=out[k])
    return (out, noises) if return_noise else out


def sdeint(sde, y0: torch.Tensor, ts: torch.Tensor, bm=None, method: Optional[str] = None, dt: float = 1e-3, adaptive: bool = False,
           options: Optional[dict] = None, names=None, seed: Optional[int] = None, **unused) -> torch.Tensor:
    """Drop-in for `torchsde.sdeint` on the reference's path: -> `[len(ts), B, D]`, row 0 = `y0`."""
    if bm is not None:
        raise NotImplementedError("a user-supplied Brownian motion is outside the reference's path (it passes none)")
    method = "euler" if method is None else method
    if method != "euler" or adaptive:
        raise NotImplementedError("the reference uses the fixed-step 'euler' scheme only")
    if getattr(sde, "noise_type", "diagonal") != "diagonal" or getattr(sde, "sde_type", "ito") != "ito":
        raise NotImplementedError("diagonal Ito noise only (what both reference models declare)")
    if not y0.is_cuda:
        raise _lib.Ab200Error("y0 must be a CUDA tensor: ananke_abm_b200 has no CPU path")
    if y0.dim() != 2 or y0.shape[1] % 4:
        raise ValueError("sdeint: y0 must be [B, D] with D a multiple of 4")
    if ts.dim() != 1 or ts.numel() < 1 or bool((ts[1:] <= ts[:-1]).any()):
        raise ValueError("ts must be one-dimensional and strictly increasing")
    params = list(sde.parameters()) if isinstance(sde, torch.nn.Module) else []
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    if torch.is_grad_enabled() and (y0.requires_grad or any(p.requires_grad for p in params)):
        # the reference's latent_ode trains THROUGH sdeint (latent_ode/train/train.py:57-74): same steps, recorded by autograd
        return _integrate_with_grad(sde, y0, ts.detach().cpu(), float(dt), int(seed))
    return _integrate(sde, y0.detach(), ts.detach().cpu(), float(dt), int(seed))
