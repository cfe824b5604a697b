"""Continuous adjoint on the tensor-core STAGE kernels for the fixed-grid rk4 solver (BASELINE.json configs[4]: "adjoint
backward, bf16 tensor-core projection"; the call shape of latent_ode/architecture/ode_components.py:50 with method='rk4').

torchdiffeq semantics (adjoint.py `OdeintAdjointMethod`, fixed_grid.py `RK4`, solvers.py `FixedGridODESolver`):
  forward   under no_grad; only the rows at the requested times are kept.  options['step_size'] = h0 makes the solver
            step over t[0] + k h0 (last point moved onto t[-1]) and interpolate the outputs linearly, so a day can be
            integrated in 96 steps while y(24) is the only row that exists: memory does not grow with the step count.
  backward  for i = T-1 .. 1 the augmented state z = [y, a_y, a_theta] is integrated from t[i] to t[i-1] with the same
            3/8-rule steps (its own step_size grid starting at t[i]), then y is re-seeded with the saved row i-1 and
            dL/dy[i-1] is added to a_y.  Augmented dynamics:  dz/dt = [ f(t, y), -a_y^T df/dy, -a_y^T df/dtheta ].

Mapping onto the stage kernels.  The drift is second order, f = [v, A(p, v, h, t), 0], so with gx = J_A^T a_v:
  -a_y^T df/dy = -[ gx.p, a_p + gx.v, gx.h ]            (one `ab200_stage_backward` launch with upstream a_v)
  -a_y^T df/dtheta = -J_theta^T a_v                      (the blobs of the same launch, folded by `ab200_wgrad_accumulate`)
and the Runge-Kutta stages of y are the usual combinations of (y0, A_1 .. A_{s-1}), independent of a_y: one fused forward launch
per step (`ab200_stage_forward_fused`), as in the forward solve.  a_theta never
feeds back into the dynamics, so its stage values are not formed: the stage's upstream gradient is scaled by
c_s = -h b_s (h < 0 going backward) and the weight-gradient accumulators ARE a_theta.  The vector-Jacobian product is
linear in its upstream, so gx comes back scaled by c_s and is divided out in the stage algebra.

Every stage's upstream gradient c_s a_v,s is LINEAR in a0 and the earlier stages' products (`fused_step_coefficients`), and so is
the step's solution a_y(t + h): an augmented step is the fused forward launch of the y stages, `ab200_pv_combine_backward` (the a0
parts of the four upstream gradients and of the solution), the vector-Jacobian products with the earlier products as gather sources
of the launch, the weight-gradient pass, and `ab200_adjoint_gather` (the solution) -- no elementwise pass between the stages.
Two launch structures (`rk4_continuous_adjoint(fused=...)`, `options["adjoint_fused"]`): the four products as ONE
`ab200_stage_backward_fused` launch (default while the blobs of four stages fit into the free memory, ~5M agents on an empty GPU), or one
`ab200_stage_backward` launch per stage with a one-stage blob ring (8M agents on one GPU).
Every evaluation of A and of its vector-Jacobian product is a tcgen05 kernel (fp16 / bf16 operands, fp32 accumulate);
y, a_y and a_theta are fp32.  Stated tolerance: that of the tensor-core path.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import stage
from .stage import RK38, TM


class _CudaLayout:
    """tile-blocked buffers on the device (library kernels)"""
    block = staticmethod(stage.rows_block)
    unblock = staticmethod(stage.rows_unblock)
    zeros = staticmethod(stage.blocked_zeros)


def step_grid(t0: float, t1: float, step_size: Optional[float], np_dtype=np.float32) -> List[float]:
    """solvers.py `_grid_constructor_from_step_size` for one solve from t0 to t1 (either direction; torchdiffeq integrates a
    decreasing interval in negated time, which mirrors the grid): t0 + sign k step, last point moved onto t1; computed in the
    time dtype like the package does.  step_size None: the default grid constructor (the interval is one step)."""
    if step_size is None:
        return [float(t0), float(t1)]
    a, b, s = np_dtype(t0), np_dtype(t1), np_dtype(step_size)
    if not s > 0:
        raise ValueError("step_size must be positive")
    sign = np_dtype(1.0) if b >= a else np_dtype(-1.0)
    niters = int(math.ceil(float((b - a) * sign / s + np_dtype(1.0))))
    g = (np.arange(0, niters).astype(np_dtype) * s) * sign + a
    g[-1] = b
    return [float(x) for x in g]


def _views(buf: torch.Tensor, F: int, P: int):
    """(p, v, h) views of a tile-blocked [Bp, F = 2P + H] buffer: float4 (tile * F/4 + f4) * 128 + row"""
    x = buf.view(-1, F // 4, TM, 4)
    return x[:, :P // 4], x[:, P // 4:2 * P // 4], x[:, 2 * P // 4:]


def rk4_forward_rows(eng, y0: torch.Tensor, t_host: Sequence[float], step_size: Optional[float] = None, lay=_CudaLayout,
                     np_dtype=np.float32) -> torch.Tensor:
    """y0 row-major [B, D] -> the rows at the requested times [T, B, D]; two state buffers whatever the number of steps."""
    B, T = y0.shape[0], len(t_host)
    D, P = eng.D, eng.P
    y_path = torch.empty((T, B, D), dtype=y0.dtype, device=y0.device)
    y_path[0].copy_(y0)
    if T == 1:
        return y_path
    grid = [float(x) for x in t_host] if step_size is None else step_grid(t_host[0], t_host[-1], step_size, np_dtype)
    yb = [lay.block(y0.contiguous()), lay.zeros(B, D, y0.device)]
    A = [lay.zeros(B, P, y0.device) for _ in range(3)]
    j = 1
    for n in range(len(grid) - 1):
        t0, t1 = grid[n], grid[n + 1]
        dt = t1 - t0
        yn, yn1 = yb[n % 2], yb[(n + 1) % 2]
        stages = [(i, RK38.stage_input(i, dt), t0 + RK38.c[i] * dt, A[i]) for i in range(3)]
        stages.append((3, RK38.stage_input(3, dt), t1, None))
        eng.stage_forward_fused(yn, A, stages, B, y_out=yn1, cout=RK38.combo(RK38.b, dt))
        while j < T and t1 >= t_host[j]:      # solvers.py FixedGridODESolver.integrate / _linear_interp
            tj = float(t_host[j])
            if tj == t1:
                lay.unblock(yn1, B, D, out=y_path[j])
            elif tj == t0:
                lay.unblock(yn, B, D, out=y_path[j])
            else:
                r0, r1 = lay.unblock(yn, B, D), lay.unblock(yn1, B, D)
                y_path[j].copy_(r0 + ((tj - t0) / (t1 - t0)) * (r1 - r0))
            j += 1
    eng.check_status()
    return y_path


class _FusedBuffers:
    def __init__(self, B: int, D: int, P: int, device, lay):
        self.A = [lay.zeros(B, P, device) for _ in range(3)]      # stage accelerations of the y part
        self.GB = [lay.zeros(B, P, device) for _ in range(4)]     # the a0 part of every stage's upstream gradient
        self.GX = [lay.zeros(B, D, device) for _ in range(4)]     # c_s J_A^T a_v,s of every stage
        self.base = lay.zeros(B, D, device)
        self.y_next = lay.zeros(B, D, device)
        self.a_next = lay.zeros(B, D, device)


def fused_step_coefficients(h: float):
    """The augmented 3/8-rule step written over (a0, gx'_0 .. gx'_3), gx'_s = c_s J_A^T a_v,s, c_s = -h b_s.  With
    ka_s = -[gx_s.p, a_p,s + gx_s.v, gx_s.h] and a_s = a0 + h sum_j beta_sj ka_j, everything a stage needs is LINEAR in a0 and
    the earlier stages' products:
        upstream_s = c_s a_v,s = cva[s] a0.v + cpa[s] a0.p + sum_{i<s} dp[s][i] gx'_i.p + dv[s][i] gx'_i.v
        a_y(t + h) = [a0.p, a0.v - h a0.p, a0.h] + sum_i [gx'_i.p, w[i] gx'_i.p + gx'_i.v, gx'_i.h]
    -- the shapes `ab200_pv_combine_backward`, `ab200_stage_backward_fused` (sources = earlier entries of the launch) and
    `ab200_adjoint_gather` already compute.  -> (c, cpa, cva, dp, dv, w)."""
    beta = [list(r) + [0.0] * (4 - len(r)) for r in RK38.beta]
    b, crk = RK38.b, RK38.c
    c = [-h * b[s] for s in range(4)]
    cva = [c[s] for s in range(4)]
    cpa = [-c[s] * h * crk[s] for s in range(4)]
    dp = [[c[s] * h * h * sum(beta[s][j] * beta[j][i] for j in range(4)) / c[i] for i in range(s)] for s in range(4)]
    dv = [[-c[s] * h * beta[s][i] / c[i] for i in range(s)] for s in range(4)]
    w = [h * h * sum(b[s] * beta[s][i] for s in range(4)) / c[i] for i in range(4)]
    return c, cpa, cva, dp, dv, w


def _aug_step(eng, yb: torch.Tensor, ab: torch.Tensor, t0: float, t1: float, B: int, w: _FusedBuffers, fused: bool):
    """One 3/8-rule step of the augmented system from t0 to t1 (h = t1 - t0, negative in the backward pass); a_theta accumulates
    inside the engine.  Returns (y(t1), a_y(t1)) as buffers of `w` swapped with the inputs.
      1. the four y stages (A_1..A_3, y(t1)): ONE fused forward launch, exactly the step of the forward solve (they do not depend on a_y)
      2. `ab200_pv_combine_backward`: base = [a0.p, a0.v - h a0.p, a0.h] and the a0 part GB[s] of every stage's upstream gradient
      3. the vector-Jacobian products GX[s] = c_s J_A^T a_v,s (+ the stage's weight-gradient blobs with the same scale); the earlier
         stages' products enter through the launch's upstream gather: `fused` = ONE `ab200_stage_backward_fused` launch whose
         entries feed each other, else one `ab200_stage_backward` launch and one weight-gradient pass per stage (one-stage blob ring)
      4. the weight-gradient pass;  5. `ab200_adjoint_gather`: a_y(t1) = base + sum_i [GX_i.p, w_i GX_i.p + GX_i.v, GX_i.h]"""
    h = t1 - t0
    cins = [RK38.stage_input(s, h) for s in range(4)]
    times = [t0 + RK38.c[s] * h for s in range(3)] + [t1]
    c, cpa, cva, dp, dv, wv = fused_step_coefficients(h)
    eng.stage_forward_fused(yb, w.A, [(s, cins[s], times[s], w.A[s] if s < 3 else None) for s in range(4)], B, y_out=w.y_next,
                            cout=RK38.combo(RK38.b, h))
    eng.combine_backward(ab, stage.Combo(-h, cpa, cva), B, w.base, w.GB, accumulate=False)
    srcs = [[i for i in range(s) if dp[s][i] != 0.0 or dv[s][i] != 0.0] for s in range(4)]
    if fused:
        eng.stage_backward_fused(yb, w.A, [(s, cins[s], times[s], w.GB[s], [(i, dp[s][i], dv[s][i]) for i in srcs[s]], w.GX[s])
                                           for s in range(4)], B)
    else:
        for s in range(4):
            eng.stage_backward(yb, w.A[:s], cins[s], times[s], B, w.GB[s], [w.GX[i] for i in srcs[s]], [dp[s][i] for i in srcs[s]],
                               [dv[s][i] for i in srcs[s]], w.GX[s])
            eng.flush()
    eng.flush()
    eng.adjoint_gather(w.base, w.GX, wv, B, w.a_next)
    y_new, a_new = w.y_next, w.a_next
    w.y_next, w.a_next = yb, ab
    return y_new, a_new


def _fused_fits(B: int, D: int, P: int, device) -> bool:
    """the fused structure keeps the blobs of FOUR stages (~3.1 KB per agent-stage) next to the 9 [Bp, D] + 7 [Bp, P] work buffers;
    it is chosen while that stays under 70 % of the memory that is free right now (device-free + the caching allocator's idle blocks),
    and under 56 GB of blobs where that cannot be asked (CPU stand-in)"""
    Bp = stage.padded_rows(B)
    blobs = Bp * 3100 * 4
    if torch.device(device).type != "cuda":
        return blobs < (56 << 30)
    free, _ = torch.cuda.mem_get_info(device)
    idle = torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
    return blobs + Bp * 4 * (9 * D + 7 * P) < 0.7 * (free + idle)


def rk4_continuous_adjoint(eng, t_host: Sequence[float], y_rows: torch.Tensor, grad_rows: torch.Tensor,
                           step_size: Optional[float] = None, lay=_CudaLayout, np_dtype=np.float32, fused: Optional[bool] = None):
    """-> (dL/dy0 row-major [B, D], a_theta(t[0]) = dL/dtheta in the drift's flat parameter order).
    `fused` (default: while the blobs of four stages fit into the free memory, `_fused_fits`: up to ~5M agents on an otherwise empty
    180 GB GPU): the four vector-Jacobian products of a step in one launch; otherwise one launch per stage with a one-stage blob ring and the weight-gradient pass after every stage
    (8M agents on one GPU)."""
    T, B, D = grad_rows.shape
    dev = grad_rows.device
    if fused is None:
        fused = _fused_fits(B, D, eng.P, dev)
    eng.backward_begin(B, 4 if fused else 1)
    w = _FusedBuffers(B, D, eng.P, dev, lay)
    ab = lay.block(grad_rows[T - 1].contiguous())
    yb = None
    for i in range(T - 1, 0, -1):
        row = y_rows[i].contiguous()
        yb = lay.block(row) if yb is None else lay.block(row, yb)      # re-seed y with the saved row (adjoint.py: aug_state[1] = y[i - 1])
        grid = step_grid(t_host[i], t_host[i - 1], step_size, np_dtype)
        for n in range(len(grid) - 1):
            yb, ab = _aug_step(eng, yb, ab, grid[n], grid[n + 1], B, w, bool(fused))
        lay.block(grad_rows[i - 1].contiguous(), ab, accumulate=True)
    gw = eng.backward_end()
    return lay.unblock(ab, B, D), gw


class _ContinuousAdjointRK4TC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, t, w_flat, spec, t_host, step_size, adj_step_size, fwd_format, fused=None):
        eng = stage.TcEngine(spec, w_flat)
        if fwd_format is not None:
            eng.fwd_format = stage.fwd_format_code(fwd_format)
        th = [float(v) for v in t_host.tolist()]
        npdt = np.float64 if t_host.dtype == torch.float64 else np.float32
        with torch.no_grad():
            y = rk4_forward_rows(eng, y0.contiguous().float(), th, step_size, np_dtype=npdt)
        ctx.eng, ctx.th, ctx.adj_step_size, ctx.npdt, ctx.fused = eng, th, adj_step_size, npdt, fused
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        (y,) = ctx.saved_tensors
        gy0, gw = rk4_continuous_adjoint(ctx.eng, ctx.th, y, grad_y.contiguous().float(), ctx.adj_step_size, np_dtype=ctx.npdt,
                                         fused=ctx.fused)
        return gy0, None, gw, None, None, None, None, None, None
