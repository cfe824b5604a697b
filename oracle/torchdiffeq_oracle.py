"""ORACLE (test infrastructure, not product code) -- CPU restatement of torchdiffeq 0.2.5.

PARITY UNPINNED for the solver arithmetic itself: the reference pins `torchdiffeq==0.2.5`
(/root/reference/uv.lock:2896-2897) but the package is neither vendored nor installed in the build
image and there is no network, and the reference ships no golden vectors for it.  What follows is a
restatement of the published algorithm (torchdiffeq 0.2.5: `_impl/fixed_grid.py`, `_impl/rk_common.py`,
`_impl/dopri5.py`, `_impl/interp.py`, `_impl/misc.py`, `_impl/solvers.py`, `_impl/adjoint.py`), anchored on
the reference's own call sites:

  * `odeint(self.odefunc, y0, times_union, method="rk4", rtol=, atol=)`
        /root/reference/src/ananke_abm/models/mode_sep/architecture/model.py:184-191
  * `odeint(self.ode_func, y0, times, method='dopri5', options={'dtype': torch.float32})`
        /root/reference/src/ananke_abm/models/latent_ode/architecture/model.py:192,196
  * `odeint_adjoint(wrapped_func, x0, t, rtol=, atol=, method='dopri5')`
        /root/reference/src/ananke_abm/models/latent_ode/architecture/ode_components.py:3,50

What pins it instead (tests/test_oracle_solver.py): the Dormand-Prince tableau is checked against
scipy.integrate.RK45's class constants, rk4 (3/8 rule) and dopri5 are checked for their convergence
order and against closed-form solutions, and the adjoint is checked against autograd through the solver.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may
import this file.  It is importable as a drop-in `torchdiffeq` module
(`sys.modules["torchdiffeq"] = oracle.torchdiffeq_oracle`) so that the *unmodified* reference model files
run on it (scripts under tests/golden/ do exactly that to mint the golden vectors).
"""
from __future__ import annotations

import warnings
from typing import Callable, List, Optional, Sequence, Tuple

import torch

__all__ = ["odeint", "odeint_adjoint"]
__version__ = "0.2.5-oracle"


# --------------------------------------------------------------------------------------------------
# misc.py
# --------------------------------------------------------------------------------------------------
def _rms_norm(x: torch.Tensor) -> torch.Tensor:
    # misc.py `_rms_norm`: sqrt(mean(x^2)) over the WHOLE tensor (all agents, all dims).
    return x.abs().pow(2).mean().sqrt()


def _mixed_norm_factory(shapes: Sequence[torch.Size]):
    # misc.py `_mixed_norm`: max over the rms norms of the tuple components (flat state is split back).
    def _norm(flat: torch.Tensor) -> torch.Tensor:
        out = []
        off = 0
        for shp in shapes:
            n = int(torch.Size(shp).numel())
            out.append(_rms_norm(flat[off:off + n]) if n > 0 else flat.new_zeros(()))
            off += n
        return torch.stack(out).max()
    return _norm


def _select_initial_step(func, t0, y0, order, rtol, atol, norm, f0):
    # misc.py `_select_initial_step` (Hairer, Norsett & Wanner I, II.4); `order` here is q = order-1.
    dtype = y0.dtype
    t_dtype = t0.dtype
    scale = atol + torch.abs(y0) * rtol
    d0 = norm(y0 / scale).abs()
    d1 = norm(f0 / scale).abs()
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = torch.tensor(1e-6, dtype=dtype, device=y0.device)
    else:
        h0 = 0.01 * d0 / d1
    h0 = h0.abs()
    y1 = y0 + h0 * f0
    f1 = func(t0 + h0, y1)
    d2 = torch.abs(norm((f1 - f0) / scale) / h0)
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6, dtype=dtype, device=y0.device), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / float(order + 1))
    h1 = h1.abs()
    return torch.min(100 * h0, h1).to(t_dtype)


def _compute_error_ratio(error_estimate, rtol, atol, y0, y1, norm):
    error_tol = atol + rtol * torch.max(y0.abs(), y1.abs())
    return norm(error_estimate / error_tol).abs()


@torch.no_grad()
def _optimal_step_size(last_step, error_ratio, safety, ifactor, dfactor, order):
    if error_ratio == 0:
        return last_step * ifactor
    if error_ratio < 1:
        dfactor = torch.ones((), dtype=last_step.dtype, device=last_step.device)
    error_ratio = error_ratio.type_as(last_step)
    exponent = torch.tensor(order, dtype=last_step.dtype, device=last_step.device).reciprocal()
    factor = torch.min(ifactor, torch.max(safety / error_ratio ** exponent, dfactor))
    return last_step * factor


class _TimeCastFunc:
    """misc.py `_PerturbFunc`: `t` is cast to `y.dtype` before every user call (perturb is off by default)."""

    def __init__(self, base):
        self.base = base

    def __call__(self, t, y):
        return self.base(t.to(y.dtype), y)


class _ReverseFunc:
    """misc.py `_ReverseFunc`: decreasing `t` is integrated as increasing `-t` with a negated field."""

    def __init__(self, base, mul: float = 1.0):
        self.base = base
        self.mul = mul

    def __call__(self, t, y):
        return self.mul * self.base(-t, y)


# --------------------------------------------------------------------------------------------------
# fixed_grid.py / rk_common.py : method="rk4" is the 3/8-rule variant, one step per interval of `t`
# --------------------------------------------------------------------------------------------------
_ONE_THIRD = 1.0 / 3.0
_TWO_THIRDS = 2.0 / 3.0


def rk4_alt_step(func, t0, dt, t1, y0):
    """rk_common.py `rk4_alt_step_func` -- same association order as the package."""
    k1 = func(t0, y0)
    k2 = func(t0 + dt * _ONE_THIRD, y0 + dt * k1 * _ONE_THIRD)
    k3 = func(t0 + dt * _TWO_THIRDS, y0 + dt * (k2 - k1 * _ONE_THIRD))
    k4 = func(t1, y0 + dt * (k1 - k2 + k3))
    return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125


def _grid_from_step_size(t, step_size):
    """solvers.py `FixedGridODESolver._grid_constructor_from_step_size`: t[0] + k * step_size, the last point moved onto t[-1]."""
    niters = int(torch.ceil((t[-1] - t[0]) / step_size + 1).item())
    grid = torch.arange(0, niters, dtype=t.dtype, device=t.device) * step_size + t[0]
    grid[-1] = t[-1]
    return grid


def _linear_interp(t0, t1, y0, y1, t):
    """solvers.py `FixedGridODESolver._linear_interp`."""
    if t == t0:
        return y0
    if t == t1:
        return y1
    slope = (t - t0) / (t1 - t0)
    return y0 + slope * (y1 - y0)


def _rk4_fixed_grid(func, y0, t, step_size=None):
    # solvers.py `FixedGridODESolver.integrate`.  Default grid constructor: grid == t, every output lands exactly on a
    # step end and `_linear_interp` returns y1 itself.  options["step_size"]: the solver steps over t[0] + k * step_size
    # and the outputs are linear interpolants between the two grid points around each requested time.
    if step_size is None:
        sol = [y0]
        y = y0
        for i in range(t.shape[0] - 1):
            t0, t1 = t[i], t[i + 1]
            dt = t1 - t0
            y = y + rk4_alt_step(func, t0, dt, t1, y)
            sol.append(y)
        return torch.stack(sol, dim=0)
    grid = _grid_from_step_size(t, step_size)
    assert grid[0] == t[0] and grid[-1] == t[-1]
    sol = [y0]
    j = 1
    y = y0
    for i in range(grid.shape[0] - 1):
        t0, t1 = grid[i], grid[i + 1]
        dt = t1 - t0
        y1 = y + rk4_alt_step(func, t0, dt, t1, y)
        while j < t.shape[0] and t1 >= t[j]:
            sol.append(_linear_interp(t0, t1, y, y1, t[j]))
            j += 1
        y = y1
    return torch.stack(sol, dim=0)


def _euler_fixed_grid(func, y0, t):
    sol = [y0]
    y = y0
    for i in range(t.shape[0] - 1):
        dt = t[i + 1] - t[i]
        y = y + dt * func(t[i], y)
        sol.append(y)
    return torch.stack(sol, dim=0)


# --------------------------------------------------------------------------------------------------
# dopri5.py / rk_common.py / interp.py
# --------------------------------------------------------------------------------------------------
DP_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
DP_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
DP_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
DP_C_ERROR = [
    35 / 384 - 1951 / 21600,
    0,
    500 / 1113 - 22642 / 50085,
    125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400,
    11 / 84 - 649 / 6300,
    -1.0 / 60.0,
]
DP_C_MID = [
    6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
    187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2,
]


def _interp_fit(y0, y1, y_mid, f0, f1, dt):
    a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
    b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
    c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
    d = dt * f0
    e = y0
    return [e, d, c, b, a]


def _interp_evaluate(coefficients, t0, t1, t):
    assert (t0 <= t) & (t <= t1), "invalid interpolation, fails `t0 <= t <= t1`"
    x = (t - t0) / (t1 - t0)
    x = x.to(coefficients[0].dtype)
    total = coefficients[0] + x * coefficients[1]
    x_power = x
    for coefficient in coefficients[2:]:
        x_power = x_power * x
        total = total + x_power * coefficient
    return total


class Dopri5Solver:
    """rk_common.py `RKAdaptiveStepsizeODESolver` specialised to the Dormand-Prince-Shampine tableau."""

    order = 5

    def __init__(self, func, y0, rtol, atol, norm=None, first_step=None, safety=0.9, ifactor=10.0,
                 dfactor=0.2, max_num_steps=2 ** 31 - 1, dtype=torch.float64, min_step=0.0,
                 max_step=float("inf"), **unused):
        if unused:
            warnings.warn(f"Dopri5Solver: Unexpected arguments {unused}")
        self.func = func
        self.y0 = y0
        dtype = torch.promote_types(dtype, y0.dtype)
        dev = y0.device
        self.dtype = dtype
        self.rtol = torch.as_tensor(rtol, dtype=dtype, device=dev)
        self.atol = torch.as_tensor(atol, dtype=dtype, device=dev)
        self.norm = _rms_norm if norm is None else norm
        self.first_step = None if first_step is None else torch.as_tensor(first_step, dtype=dtype, device=dev)
        self.safety = torch.as_tensor(safety, dtype=dtype, device=dev)
        self.ifactor = torch.as_tensor(ifactor, dtype=dtype, device=dev)
        self.dfactor = torch.as_tensor(dfactor, dtype=dtype, device=dev)
        self.min_step = torch.as_tensor(min_step, dtype=dtype, device=dev)
        self.max_step = torch.as_tensor(max_step, dtype=dtype, device=dev)
        self.max_num_steps = int(max_num_steps)
        yd = y0.dtype
        self.alpha = [torch.tensor(a, dtype=torch.float64).to(yd).to(dev) for a in DP_ALPHA]
        self.beta = [torch.tensor(b, dtype=torch.float64).to(yd).to(dev) for b in DP_BETA]
        self.c_sol = torch.tensor(DP_C_SOL, dtype=torch.float64).to(yd).to(dev)
        self.c_error = torch.tensor(DP_C_ERROR, dtype=torch.float64).to(yd).to(dev)
        self.mid = torch.tensor(DP_C_MID, dtype=torch.float64).to(yd).to(dev)
        # statistics for tests / bench (not part of the package)
        self.n_accepted = 0
        self.n_rejected = 0
        self.step_log: List[Tuple[float, float, bool]] = []

    # -- one attempted step: rk_common.py `_runge_kutta_step`
    def _rk_step(self, y0, f0, t0, dt, t1):
        t0y = t0.to(y0.dtype)
        dty = dt.to(y0.dtype)
        t1y = t1.to(y0.dtype)
        ks = [f0]
        yi = y0
        for i, (alpha_i, beta_i) in enumerate(zip(self.alpha, self.beta)):
            ti = t1y if float(DP_ALPHA[i]) == 1.0 else t0y + alpha_i * dty
            kstack = torch.stack(ks, dim=-1)                      # [..., i+1]
            yi = y0 + torch.sum(kstack * (beta_i * dty), dim=-1).view_as(f0)
            ks.append(self.func(ti, yi))
        k = torch.stack(ks, dim=-1)                               # [..., 7]
        # FSAL tableau: c_sol[-1]==0 and c_sol[:-1]==beta[-1], so y1 is the last stage input
        y1 = yi
        f1 = ks[-1]
        y1_error = k.matmul(dty * self.c_error)
        return y1, f1, y1_error, k

    def _before_integrate(self, t):
        f0 = self.func(t[0], self.y0)
        if self.first_step is None:
            first_step = _select_initial_step(self.func, t[0], self.y0, self.order - 1, self.rtol, self.atol,
                                              self.norm, f0=f0)
        else:
            first_step = self.first_step
        self.state = (self.y0, f0, t[0], t[0], first_step, [self.y0] * 5)

    def _adaptive_step(self, state):
        y0, f0, _, t0, dt, interp_coeff = state
        t1 = t0 + dt
        assert t0 + dt > t0, "underflow in dt {}".format(dt.item())
        assert torch.isfinite(y0).all(), "non-finite values in state `y`: {}".format(y0)
        y1, f1, y1_error, k = self._rk_step(y0, f0, t0, dt, t1)
        error_ratio = _compute_error_ratio(y1_error, self.rtol, self.atol, y0, y1, self.norm)
        accept_step = bool(error_ratio <= 1)
        if dt > self.max_step:
            accept_step = False
        if dt <= self.min_step:
            accept_step = True
        self.step_log.append((float(t0.detach()), float(dt.detach()), accept_step))
        if accept_step:
            self.n_accepted += 1
            t_next, y_next, f_next = t1, y1, f1
            dty = dt.type_as(y0)
            y_mid = y0 + k.matmul(dty * self.mid).view_as(y0)
            interp_coeff = _interp_fit(y0, y1, y_mid, k[..., 0], k[..., -1], dty)
        else:
            self.n_rejected += 1
            t_next, y_next, f_next = t0, y0, f0
        dt_next = _optimal_step_size(dt, error_ratio, self.safety, self.ifactor, self.dfactor, self.order)
        dt_next = dt_next.clamp(self.min_step, self.max_step)
        return (y_next, f_next, t0, t_next, dt_next, interp_coeff)

    def _advance(self, next_t):
        n_steps = 0
        while next_t > self.state[3]:
            assert n_steps < self.max_num_steps, "max_num_steps exceeded ({}>={})".format(n_steps, self.max_num_steps)
            self.state = self._adaptive_step(self.state)
            n_steps += 1
        return _interp_evaluate(self.state[5], self.state[2], self.state[3], next_t)

    def integrate(self, t):
        sol = [self.y0]
        t = t.to(self.dtype)
        self._before_integrate(t)
        for i in range(1, len(t)):
            sol.append(self._advance(t[i]))
        return torch.stack(sol, dim=0)


# --------------------------------------------------------------------------------------------------
# odeint.py
# --------------------------------------------------------------------------------------------------
_LAST_SOLVER = {"solver": None}   # test hook: the most recent adaptive solver (step log, counters)


def _check_t(t: torch.Tensor):
    assert isinstance(t, torch.Tensor), "t must be a torch.Tensor"
    assert t.ndimension() == 1, "t must be one dimensional"
    assert torch.is_floating_point(t), "t must be a floating point Tensor"
    d = t[1:] - t[:-1]
    if not (bool((d > 0).all()) or bool((d < 0).all())):
        raise AssertionError("t must be strictly increasing or decreasing")


def odeint(func: Callable, y0, t: torch.Tensor, *, rtol: float = 1e-7, atol: float = 1e-9,
           method: Optional[str] = None, options: Optional[dict] = None, event_fn=None):
    """odeint.py `odeint`: returns `[len(t), *y0.shape]`, row 0 is `y0`.  Tuple states are flattened."""
    if event_fn is not None:
        raise NotImplementedError("event handling is not on the reference's path")
    options = {} if options is None else dict(options)
    method = "dopri5" if method is None else method
    _check_t(t)

    is_tuple = isinstance(y0, (tuple, list))
    if is_tuple:
        shapes = [y.shape for y in y0]
        flat0 = torch.cat([y.reshape(-1) for y in y0])
        user = func

        def func_flat(tt, yy):
            parts, off = [], 0
            for shp in shapes:
                n = int(torch.Size(shp).numel())
                parts.append(yy[off:off + n].view(shp))
                off += n
            out = user(tt, tuple(parts))
            return torch.cat([o.reshape(-1) for o in out])
        func, y0 = func_flat, flat0
        if "norm" not in options:
            options["norm"] = _mixed_norm_factory(shapes)

    if t.numel() > 1 and bool(t[0] > t[1]):
        t = -t
        func = _ReverseFunc(func, mul=-1.0)
    f = _TimeCastFunc(func)

    if method == "rk4":
        options.pop("norm", None)
        options.pop("dtype", None)
        sol = _rk4_fixed_grid(f, y0, t, options.pop("step_size", None))
    elif method == "euler":
        options.pop("norm", None)
        sol = _euler_fixed_grid(f, y0, t)
    elif method == "dopri5":
        solver = Dopri5Solver(f, y0, rtol=rtol, atol=atol, **options)
        sol = solver.integrate(t)
        _LAST_SOLVER["solver"] = solver
    else:
        raise ValueError(f'Invalid method "{method}".')

    if is_tuple:
        outs, off = [], 0
        for shp in shapes:
            n = int(torch.Size(shp).numel())
            outs.append(sol[:, off:off + n].view(len(t), *shp))
            off += n
        return tuple(outs)
    return sol


# --------------------------------------------------------------------------------------------------
# adjoint.py
# --------------------------------------------------------------------------------------------------
class _OdeintAdjoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, func, y0, t, rtol, atol, method, options, adjoint_rtol, adjoint_atol, adjoint_method,
                adjoint_options, *adjoint_params):
        ctx.func = func
        ctx.cfg = (adjoint_rtol, adjoint_atol, adjoint_method, adjoint_options)
        with torch.no_grad():
            y = odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=options)
        ctx.save_for_backward(t, y, *adjoint_params)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        func = ctx.func
        adjoint_rtol, adjoint_atol, adjoint_method, adjoint_options = ctx.cfg
        t, y, *adjoint_params = ctx.saved_tensors
        adjoint_params = tuple(adjoint_params)
        with torch.no_grad():
            aug_state = [torch.zeros((), dtype=y.dtype, device=y.device), y[-1], grad_y[-1]]
            aug_state.extend([torch.zeros_like(p) for p in adjoint_params])

            def augmented_dynamics(tt, y_aug):
                yy = y_aug[1]
                adj_y = y_aug[2]
                with torch.enable_grad():
                    t_ = tt.detach()
                    yy = yy.detach().requires_grad_(True)
                    func_eval = func(t_, yy)
                    vjps = torch.autograd.grad(func_eval, (yy,) + adjoint_params, -adj_y,
                                               allow_unused=True, retain_graph=True)
                vjp_y, *vjp_params = vjps
                vjp_y = torch.zeros_like(yy) if vjp_y is None else vjp_y
                vjp_params = [torch.zeros_like(p) if v is None else v for p, v in zip(adjoint_params, vjp_params)]
                return (torch.zeros_like(t_), func_eval, vjp_y, *vjp_params)

            # adjoint.py `handle_adjoint_norm_`: "seminorm" = the default mixed norm without the parameter adjoints
            adjoint_options = dict(adjoint_options)
            if adjoint_options.get("norm") == "seminorm":
                keep = _mixed_norm_factory([a.shape for a in aug_state[:3]])      # (t, y, a_y) lead the flattened state
                adjoint_options["norm"] = keep

            for i in range(len(t) - 1, 0, -1):
                sol = odeint(augmented_dynamics, tuple(aug_state), t[i - 1:i + 1].flip(0),
                             rtol=adjoint_rtol, atol=adjoint_atol, method=adjoint_method,
                             options=dict(adjoint_options))
                aug_state = [a[1] for a in sol]
                aug_state[1] = y[i - 1]
                aug_state[2] = aug_state[2] + grad_y[i - 1]
            adj_y = aug_state[2]
            adj_params = aug_state[3:]
        return (None, adj_y, None, None, None, None, None, None, None, None, None, *adj_params)


def odeint_adjoint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None,
                   adjoint_rtol=None, adjoint_atol=None, adjoint_method=None, adjoint_options=None,
                   adjoint_params=None):
    """adjoint.py `odeint_adjoint`: forward under no_grad, backward by integrating the augmented system
    [vjp_t, y, a_y, a_theta...] from t[i] to t[i-1], re-seeding y with the saved forward value."""
    if event_fn is not None:
        raise NotImplementedError
    if adjoint_params is None and not isinstance(func, torch.nn.Module):
        raise ValueError("func must be an instance of nn.Module to specify the adjoint parameters; "
                         "alternatively they can be specified explicitly via the `adjoint_params` argument.")
    adjoint_rtol = rtol if adjoint_rtol is None else adjoint_rtol
    adjoint_atol = atol if adjoint_atol is None else adjoint_atol
    adjoint_method = method if adjoint_method is None else adjoint_method
    if adjoint_options is None:
        adjoint_options = {k: v for k, v in options.items() if k != "norm"} if options is not None else {}
    else:
        adjoint_options = dict(adjoint_options)
    if adjoint_params is None:
        adjoint_params = tuple(p for p in func.parameters() if p.requires_grad)
    else:
        adjoint_params = tuple(p for p in adjoint_params if p.requires_grad)
    return _OdeintAdjoint.apply(func, y0, t, rtol, atol, method, options, adjoint_rtol, adjoint_atol,
                                adjoint_method, adjoint_options, *adjoint_params)
