"""Host logic of the generic adaptive driver (`ananke_abm_b200.odeint._Dopri5`) checked on the CPU against the oracle solver.

The driver's arithmetic lives in two C-ABI kernels (`ab200_rk_stage_combine`, `ab200_rk_combine_errnorm`); here they are replaced
by numpy stand-ins that follow the header's contract, so that the CONTROL logic -- step acceptance, dense output, torchdiffeq's
mixed norm over packed components, and the gradient of autograd through the solver ops -- can be compared with oracle/torchdiffeq_oracle.py without a GPU.  (The kernels themselves are tested on the GPU.)
"""
import ctypes as C
import importlib

import numpy as np
import pytest
import torch
from torch import nn

from oracle import torchdiffeq_oracle as tdq


def _arr(ptr, n):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n,))


class _FakeLib:
    """numpy stand-ins of the two generic-path kernels (include/ananke_b200.h)"""

    def ab200_rk_combine_errnorm(self, y0, kptrs, csol, cerr, n_k, dt, rtol, atol, y1_out, sumsq, n, stream):
        ks = C.cast(kptrs, C.POINTER(C.c_void_p))
        cs, ce = C.cast(csol, C.POINTER(C.c_float)), C.cast(cerr, C.POINTER(C.c_float))
        a = _arr(y0, n).astype(np.float32)
        s = np.zeros(n, np.float32)
        e = np.zeros(n, np.float32)
        for j in range(n_k):
            k = _arr(ks[j], n)
            s = s + k * np.float32(np.float32(cs[j]) * np.float32(dt))
            e = e + k * np.float32(np.float32(ce[j]) * np.float32(dt))
        b = a + s
        tol = np.float32(atol) + np.float32(rtol) * np.maximum(np.abs(a), np.abs(b))
        _arr(sumsq, 1)[0] += np.float32(((e / tol).astype(np.float64) ** 2).sum())
        return 0


def _fake_combine(y, ks, coef, dt):
    acc = torch.zeros_like(y)
    for k, c in zip(ks, coef):
        acc = acc + k * float(torch.tensor(float(c), dtype=torch.float32) * torch.tensor(float(dt), dtype=torch.float32))
    return y + acc


@pytest.fixture()
def host_driver(monkeypatch):
    oi = importlib.import_module("ananke_abm_b200.odeint")
    from ananke_abm_b200 import _lib
    monkeypatch.setattr(_lib, "lib", lambda: _FakeLib())
    monkeypatch.setattr(oi, "_combine", _fake_combine)
    monkeypatch.setattr(oi, "_stream_ptr", lambda: None)
    return oi


class _Func(nn.Module):
    def __init__(self, dim=6):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim + 2, 16), nn.Tanh(), nn.Linear(16, dim))

    def forward(self, t, y):
        te = torch.stack([torch.sin(t * 0.7), torch.cos(t * 0.7)]).to(y.dtype).expand(y.shape[0], 2)
        return self.net(torch.cat([y, te], dim=-1))


@pytest.mark.parametrize("time_dtype", [torch.float64, torch.float32])
def test_generic_dopri5_matches_oracle_values_steps_and_autograd(host_driver, time_dtype):
    """same accepted / rejected sequence, same dense output, and the same gradient autograd produces through torchdiffeq's solver
    ops.  torchdiffeq's initial step-size heuristic runs with autograd on (misc.py `_select_initial_step`; every later step size
    comes from the no_grad `_optimal_step_size`), so dt_0 formally depends on (y0, f0); the driver treats it as a constant because a
    converged solution does not depend on its step sizes beyond the tolerance: the oracle's own gradient with the first step frozen
    (`first_step=`) agrees with its free-running gradient to round-off (asserted below)."""
    oi = host_driver
    torch.manual_seed(0)
    f = _Func()
    y0 = torch.randn(5, 6)
    t = torch.tensor([0.0, 0.4, 1.1, 2.0])
    ya = y0.clone().requires_grad_(True)
    ref = tdq.odeint(f, ya, t, method="dopri5", rtol=1e-4, atol=1e-5, options={"dtype": time_dtype})
    log = tdq._LAST_SOLVER["solver"].step_log
    ref.square().mean().backward()
    g_ref = [ya.grad.clone()] + [p.grad.clone() for p in f.parameters()]
    f.zero_grad()
    yb = y0.clone().requires_grad_(True)
    solver = oi._Dopri5(lambda tt, yy: f(tt.to(yy.dtype), yy), yb, 1e-4, 1e-5, dtype=time_dtype)
    out = solver.integrate(t)
    out.square().mean().backward()
    assert (solver.n_accepted, solver.n_rejected) == (sum(1 for x in log if x[2]), sum(1 for x in log if not x[2]))
    assert float((out - ref).abs().max()) < 2e-6 * float(ref.abs().max())
    g = [yb.grad] + [p.grad for p in f.parameters()]
    for a, b in zip(g, g_ref):
        assert float((a - b).abs().max()) < 2e-5 * float(b.abs().max()) + 1e-9
    # the oracle itself: gradient with dt_0 frozen == gradient with autograd through the step-size heuristic
    f.zero_grad()
    yc = y0.clone().requires_grad_(True)
    tdq.odeint(f, yc, t, method="dopri5", rtol=1e-4, atol=1e-5, options={"dtype": time_dtype, "first_step": float(log[0][1])}).square().mean().backward()
    assert float((yc.grad - g_ref[0]).abs().max()) < 1e-5 * float(g_ref[0].abs().max())


def test_mixed_norm_over_packed_components_matches_oracle_tuple_state(host_driver):
    """torchdiffeq's `_mixed_norm` (max over per-component RMS) on a packed state == the oracle's tuple-state solve"""
    oi = host_driver
    torch.manual_seed(1)
    A, Bm = torch.randn(4, 4) * 0.5, torch.randn(3, 3) * 2.0

    def f_tuple(t, ys):
        return (ys[0] @ A.T * torch.cos(t), torch.tanh(ys[1] @ Bm.T) * 30.0)

    y = (torch.randn(7, 4), torch.randn(2, 3) * 1e-3)
    t = torch.tensor([0.0, 0.5, 1.5])
    ref = tdq.odeint(f_tuple, y, t, method="dopri5", rtol=1e-5, atol=1e-6)
    log = tdq._LAST_SOLVER["solver"].step_log
    n0, n1 = y[0].numel(), y[1].numel()
    o1 = (n0 + 3) // 4 * 4

    def f_flat(tt, z):
        a, b = f_tuple(tt, (z[:n0].view(7, 4), z[o1:o1 + n1].view(2, 3)))
        out = torch.zeros_like(z)
        out[:n0], out[o1:o1 + n1] = a.reshape(-1), b.reshape(-1)
        return out

    z0 = torch.zeros(o1 + n1)
    z0[:n0], z0[o1:o1 + n1] = y[0].reshape(-1), y[1].reshape(-1)
    with torch.no_grad():
        mixed = oi._Dopri5(f_flat, z0, 1e-5, 1e-6, segments=[(0, n0), (o1, n1)])
        out = mixed.integrate(t)
        plain = oi._Dopri5(f_flat, z0, 1e-5, 1e-6)
        plain.integrate(t)
    assert (mixed.n_accepted, mixed.n_rejected) == (sum(1 for x in log if x[2]), sum(1 for x in log if not x[2]))
    assert float((out[:, :n0].view(3, 7, 4) - ref[0]).abs().max()) < 5e-6 * float(ref[0].abs().max())
    assert float((out[:, o1:o1 + n1].view(3, 2, 3) - ref[1]).abs().max()) < 5e-6 * float(ref[1].abs().max())
    assert plain.n_accepted != mixed.n_accepted or plain.n_rejected != mixed.n_rejected     # one RMS over everything is a different controller
