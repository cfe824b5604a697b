"""Continuous adjoint (torchdiffeq `odeint_adjoint` semantics, adjoint.py) on the GPU.

Forward under no_grad keeps only the requested output rows; backward integrates the augmented system
[y, a_y, a_theta...] from t[i] to t[i-1], re-seeding y with the saved row and adding dL/dy[i-1] to a_y -- the
O(1)-in-steps memory scheme of /root/reference/src/ananke_abm/models/latent_ode/architecture/ode_components.py:50.

For the two reference drift shapes (`drift.describe_drift`) the augmented dynamics are evaluated by kernels only:
`ab200_drift_eval` for f and `ab200_drift_vjp` for (-a_y^T df/dy, -a_y^T df/dtheta); any other `func` is evaluated by
the caller's torch code and differentiated with autograd, as torchdiffeq does.  Stage algebra and error norms run through
the fused elementwise kernels (`odeint`'s generic path) with torchdiffeq's MIXED norm: the accepted error is the max over
the RMS norms of the components (y, a_y and each parameter's adjoint), not one RMS over the packed vector.
"""
from __future__ import annotations

import torch


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


class _ContinuousAdjoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, func, y0, t, cfg, *params):
        from .odeint import odeint
        rtol, atol, method, options = cfg["fwd"]
        with torch.no_grad():
            y = odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=options)
        ctx.func, ctx.cfg = func, cfg
        ctx.save_for_backward(t, y, *params)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        from .odeint import odeint, drift_eval, drift_vjp
        from .drift import describe_drift
        func = ctx.func
        rtol, atol, method, options = ctx.cfg["adj"]
        t, y, *params = ctx.saved_tensors
        params = tuple(params)
        shape = y.shape[1:]
        n = int(y[0].numel())
        sizes = [int(p.numel()) for p in params]
        # packed state: [y | a_y | a_theta_1 | a_theta_2 ...], every component starting at a multiple of 4 floats (128-bit kernels)
        offs, off = [], 0
        for m in [n, n] + sizes:
            offs.append(off)
            off += _pad4(m)
        total = off
        segments = [(o, m) for o, m in zip(offs, [n, n] + sizes)]

        spec = describe_drift(func) if len(shape) == 2 and y.dtype == torch.float32 else None
        if spec is not None:
            sp = {id(p) for p in spec.params}
            if shape[1] != spec.state_dim or any(id(p) not in sp for p in params):
                spec = None      # adjoint_params outside the drift net: differentiate with autograd

        def pack(parts):
            z = torch.zeros(total, dtype=y.dtype, device=y.device)
            for (o, m), v in zip(segments, parts):
                z[o:o + m] = v.reshape(-1)
            return z

        if spec is not None:
            w_flat = spec.flat_params().detach()
            # position of every adjoint parameter inside the drift's flat parameter vector
            where, pos = {}, 0
            for p in spec.params:
                where[id(p)] = pos
                pos += p.numel()

            def aug(tt, z):
                tf = float(tt)
                yy = z[:n].view(shape)
                a_y = z[offs[1]:offs[1] + n].view(shape)
                f = drift_eval(spec, w_flat, tf, yy)
                gy, gw = drift_vjp(spec, w_flat, tf, yy, -a_y)
                return pack([f, gy] + [gw[where[id(p)]:where[id(p)] + p.numel()] for p in params])
        else:
            def aug(tt, z):
                yy = z[:n].view(shape)
                a_y = z[offs[1]:offs[1] + n].view(shape)
                with torch.enable_grad():
                    yy = yy.detach().requires_grad_(True)
                    f = func(tt.detach() if torch.is_tensor(tt) else torch.tensor(tt, dtype=y.dtype, device=y.device), yy)
                    vj = torch.autograd.grad(f, (yy,) + params, -a_y, allow_unused=True, retain_graph=False)
                vy = torch.zeros_like(yy) if vj[0] is None else vj[0]
                vp = [torch.zeros_like(p) if v is None else v for p, v in zip(params, vj[1:])]
                return pack([f.detach(), vy] + vp)

        adj_options = dict(options or {})
        for k_ in ("precision", "error_norm", "forward_operands", "fp16_forward", "group", "adjoint_mode", "saved_operands"):
            adj_options.pop(k_, None)      # forward-solve switches of the tensor-core path mean nothing to the augmented system
        # torchdiffeq adjoint.py `handle_adjoint_norm_`: adjoint_options["norm"] == "seminorm" keeps the parameter adjoints out of
        # the accepted-error norm (they are integrals, nothing depends on them: Kidger et al., "Hey, that's not an ODE"); the
        # default is the mixed norm over every component.  A callable norm is not supported by the fused error-norm kernel.
        adj_norm = adj_options.pop("norm", None)
        if adj_norm is not None and adj_norm != "seminorm":
            raise NotImplementedError("adjoint_options['norm']: only 'seminorm' (or nothing: torchdiffeq's default mixed norm) is supported")
        if method == "dopri5":
            adj_options["segments"] = segments[:2] if adj_norm == "seminorm" else segments
            if spec is not None:
                adj_options["time_as_float"] = True
        with torch.no_grad():
            a_y = grad_y[-1].reshape(-1).clone()
            a_p = [torch.zeros(s, dtype=y.dtype, device=y.device) for s in sizes]
            for i in range(t.numel() - 1, 0, -1):
                z = pack([y[i], a_y] + a_p)
                sol = odeint(aug, z, t[i - 1:i + 1].flip(0), rtol=rtol, atol=atol, method=method, options=dict(adj_options))
                z1 = sol[1]
                a_y = z1[offs[1]:offs[1] + n] + grad_y[i - 1].reshape(-1)
                a_p = [z1[o:o + m] for (o, m) in segments[2:]]
            outs = [a.view_as(p) for a, p in zip(a_p, params)]
        return (None, a_y.view(shape), None, None, *outs)


def _tensor_core_rk4(func, y0, t, method, options, adjoint_method, adjoint_options, adjoint_params):
    """The fixed-grid rk4 solve of a recognised drift in the tensor-core mode (options['precision'] = 'bf16'): forward and
    augmented backward system on the tcgen05 stage kernels (adjoint_tc.py).  -> the solution, or None when the call is not of
    that form (other solver, strict fp32, a generic func, hand-picked adjoint_params, decreasing t)."""
    from .odeint import _DEFAULT_PRECISION, _check_t, _require_cuda
    from .drift import describe_drift
    opts = dict(options or {})
    if method != "rk4" or (adjoint_method is not None and adjoint_method != "rk4"):
        return None
    if opts.get("precision", _DEFAULT_PRECISION["value"]) != "bf16" or not torch.is_tensor(y0) or y0.dim() != 2 or y0.dtype != torch.float32:
        return None
    spec = describe_drift(func)
    if spec is None or not spec.tc_stage_supported() or y0.shape[1] != spec.state_dim or y0.shape[0] == 0:
        return None
    if adjoint_params is not None:
        want, have = {id(p) for p in adjoint_params if p.requires_grad}, {id(p) for p in spec.params}
        if want != have:
            return None
    _require_cuda(y0, "y0")
    t_host = _check_t(t)
    if t_host.numel() > 1 and bool(t_host[0] > t_host[1]):
        return None
    step = opts.get("step_size")
    adj_step = step if adjoint_options is None else dict(adjoint_options).get("step_size")
    known = {"precision", "step_size", "forward_operands", "adjoint_fused", "adjoint_mode", "saved_operands", "dtype", "norm",
             "time_as_float", "segments"}
    extra = sorted((set(opts) | set(dict(adjoint_options or {}))) - known)
    if extra:      # torchdiffeq warns about options a solver does not use and carries on
        import warnings
        warnings.warn(f"rk4: Unexpected arguments {extra}")
    from .adjoint_tc import _ContinuousAdjointRK4TC
    return _ContinuousAdjointRK4TC.apply(y0, t, spec.flat_params(), spec, t_host, None if step is None else float(step),
                                         None if adj_step is None else float(adj_step), opts.get("forward_operands"),
                                         opts.get("adjoint_fused"))


def continuous_adjoint(func, y0, t, *, rtol, atol, method, options, adjoint_rtol=None, adjoint_atol=None,
                       adjoint_method=None, adjoint_options=None, adjoint_params=None):
    out = _tensor_core_rk4(func, y0, t, method, options, adjoint_method, adjoint_options, adjoint_params)
    if out is not None:
        return out
    if adjoint_params is None:
        adjoint_params = tuple(p for p in func.parameters() if p.requires_grad)
    else:
        adjoint_params = tuple(p for p in adjoint_params if p.requires_grad)
    if adjoint_options is None:
        adjoint_options = {k: v for k, v in (options or {}).items() if k != "norm"}
    cfg = {"fwd": (rtol, atol, method, options),
           "adj": (rtol if adjoint_rtol is None else adjoint_rtol, atol if adjoint_atol is None else adjoint_atol,
                   method if adjoint_method is None else adjoint_method, adjoint_options)}
    return _ContinuousAdjoint.apply(func, y0, t, cfg, *adjoint_params)
