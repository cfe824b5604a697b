"""Pins for the restated torchdiffeq-0.2.5 solver (oracle/torchdiffeq_oracle.py).

The real package is absent, so the pins are: (1) the Dormand-Prince tableau against
scipy.integrate.RK45's constants; (2) convergence order of rk4 (3/8 rule) and of dopri5's 5th/4th-order
pair; (3) closed-form solutions; (4) adjoint gradients against autograd through the solver.
"""
import math

import numpy as np
import pytest
import torch

from oracle import torchdiffeq_oracle as tdq


def test_dopri5_tableau_matches_scipy():
    from scipy.integrate import RK45
    # scipy: C (nodes), A (lower-triangular stage matrix), B (5th-order weights), E (error weights, B - B*)
    assert np.allclose(RK45.C[1:], tdq.DP_ALPHA[:5], rtol=0, atol=1e-15)
    for i, row in enumerate(tdq.DP_BETA[:5]):
        assert np.allclose(RK45.A[i + 1, : i + 1], row, rtol=0, atol=1e-15)
    assert np.allclose(RK45.B, tdq.DP_C_SOL[:6], rtol=0, atol=1e-15)
    assert np.allclose(tdq.DP_BETA[5], tdq.DP_C_SOL[:6])
    # scipy's E (7 entries, FSAL stage last) is B - B_hat for the classic 4th-order companion; the
    # Shampine variant torchdiffeq uses (1951/21600, 22642/50085, ...) is exactly -2/3 of it, i.e. the
    # same embedded estimator direction with a smaller constant.
    assert np.allclose(np.asarray(RK45.E) * (-2.0 / 3.0), tdq.DP_C_ERROR, rtol=0, atol=1e-15)
    assert abs(sum(tdq.DP_C_ERROR)) < 1e-15
    assert abs(sum(tdq.DP_C_MID) - 0.5) < 1e-12      # mid-point weights integrate a constant field to dt/2


def _osc(t, y):           # harmonic oscillator with unit frequency, y=[x, x']
    return torch.stack([y[..., 1], -y[..., 0]], dim=-1)


def test_rk4_is_three_eighths_rule_and_fourth_order():
    y0 = torch.tensor([[1.0, 0.0]], dtype=torch.float64)
    errs = []
    for n in (16, 32, 64):
        t = torch.linspace(0, 2.0, n + 1, dtype=torch.float64)
        y = tdq.odeint(_osc, y0, t, method="rk4")
        exact = torch.tensor([[math.cos(2.0), -math.sin(2.0)]], dtype=torch.float64)
        errs.append(float((y[-1] - exact).abs().max()))
    orders = [math.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert all(3.8 < o < 4.2 for o in orders), orders
    # one step of y'=y: the 3/8 rule and the classic rule give the same polynomial 1+h+h^2/2+h^3/6+h^4/24
    h = 0.1
    y = tdq.odeint(lambda t, y: y, torch.tensor([1.0], dtype=torch.float64), torch.tensor([0.0, h], dtype=torch.float64),
                   method="rk4")
    assert abs(float(y[-1]) - (1 + h + h ** 2 / 2 + h ** 3 / 6 + h ** 4 / 24)) < 1e-15
    # ...but differ for a t-dependent field; check the 3/8 nodes (t0+h/3, t0+2h/3) explicitly on y'=t^3
    y = tdq.odeint(lambda t, y: (t ** 3).expand_as(y), torch.tensor([0.0], dtype=torch.float64),
                   torch.tensor([0.0, 1.0], dtype=torch.float64), method="rk4")
    three_eighths = (0 + 3 * (1 / 27) + 3 * (8 / 27) + 1) / 8
    assert abs(float(y[-1]) - three_eighths) < 1e-15


def test_rk4_output_rows_and_time_cast():
    seen = []

    def f(t, y):
        seen.append(t.dtype)
        return -y
    y0 = torch.ones(3, 2, dtype=torch.float32)
    t = torch.tensor([0.0, 0.5, 0.75, 2.0], dtype=torch.float64)
    y = tdq.odeint(f, y0, t, method="rk4", rtol=1e-5, atol=1e-5)
    assert y.shape == (4, 3, 2) and y.dtype == torch.float32
    assert torch.equal(y[0], y0)
    assert all(d == torch.float32 for d in seen) and len(seen) == 12


def test_non_monotone_t_raises():
    with pytest.raises(AssertionError):
        tdq.odeint(lambda t, y: -y, torch.ones(1), torch.tensor([0.0, 1.0, 0.5]), method="rk4")


def test_decreasing_t():
    y0 = torch.tensor([[1.0, 0.0]], dtype=torch.float64)
    t = torch.linspace(0.0, -1.0, 33, dtype=torch.float64)
    y = tdq.odeint(_osc, y0, t, method="rk4")
    assert abs(float(y[-1, 0, 0]) - math.cos(-1.0)) < 1e-7
    assert abs(float(y[-1, 0, 1]) + math.sin(-1.0)) < 1e-7


def test_dopri5_closed_form_and_tolerance_scaling():
    y0 = torch.tensor([[1.0, 0.0]], dtype=torch.float64)
    t = torch.tensor([0.0, 0.3, 1.7, 5.0], dtype=torch.float64)
    exact = torch.stack([torch.cos(t), -torch.sin(t)], dim=-1).unsqueeze(1)
    errs = []
    for tol in (1e-4, 1e-6, 1e-8):
        y = tdq.odeint(_osc, y0, t, method="dopri5", rtol=tol, atol=tol)
        errs.append(float((y - exact).abs().max()))
    assert errs[0] < 5e-3 and errs[1] < 5e-5 and errs[2] < 5e-7
    assert errs[0] > errs[1] > errs[2]


def test_dopri5_pair_orders():
    """Local error of the propagated solution is O(h^6); the embedded error estimate is O(h^5)."""
    y0 = torch.tensor([1.0, 0.0], dtype=torch.float64)
    sol_err, est = [], []
    for h in (0.2, 0.1, 0.05):
        s = tdq.Dopri5Solver(_osc, y0, rtol=1e-9, atol=1e-9)
        hh = torch.tensor(h, dtype=torch.float64)
        t0 = torch.tensor(0.0, dtype=torch.float64)
        y1, f1, y1_err, k = s._rk_step(y0, _osc(t0, y0), t0, hh, t0 + hh)
        exact = torch.tensor([math.cos(h), -math.sin(h)], dtype=torch.float64)
        sol_err.append(float((y1 - exact).abs().max()))
        est.append(float(y1_err.abs().max()))
    o_sol = [math.log2(sol_err[i] / sol_err[i + 1]) for i in range(2)]
    o_est = [math.log2(est[i] / est[i + 1]) for i in range(2)]
    assert all(5.6 < o < 6.4 for o in o_sol), o_sol
    assert all(4.6 < o < 5.4 for o in o_est), o_est


def test_dopri5_dense_output_is_fourth_order_accurate():
    y0 = torch.tensor([1.0, 0.0], dtype=torch.float64)
    t = torch.linspace(0, 3.0, 61, dtype=torch.float64)         # outputs much denser than the steps
    y = tdq.odeint(_osc, y0, t, method="dopri5", rtol=1e-7, atol=1e-9)
    exact = torch.stack([torch.cos(t), -torch.sin(t)], dim=-1)
    assert float((y - exact).abs().max()) < 2e-6
    s = tdq._LAST_SOLVER["solver"]
    assert s.n_accepted < 60                                     # i.e. the interpolant was really used


def test_dopri5_float32_time_option():
    y0 = torch.tensor([[1.0, 0.0]], dtype=torch.float32)
    t = torch.tensor([0.0, 1.0, 2.0], dtype=torch.float32)
    y = tdq.odeint(_osc, y0, t, method="dopri5", rtol=1e-5, atol=1e-5, options={"dtype": torch.float32})
    s = tdq._LAST_SOLVER["solver"]
    assert s.dtype == torch.float32 and y.dtype == torch.float32
    assert abs(float(y[-1, 0, 0]) - math.cos(2.0)) < 1e-3


class _Lin(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Linear(3, 3).double()

    def forward(self, t, y):
        return torch.tanh(self.a(y)) * torch.cos(t)


@pytest.mark.parametrize("method", ["rk4", "dopri5"])
def test_adjoint_matches_autograd_through_solver(method):
    f = _Lin()
    y0 = torch.tensor([[0.3, -0.2, 0.5], [0.1, 0.4, -0.6]], dtype=torch.float64, requires_grad=True)
    t = torch.linspace(0, 1.0, 9 if method == "rk4" else 4, dtype=torch.float64)
    kw = dict(method=method, rtol=1e-10, atol=1e-10)
    w = torch.arange(1, 1 + t.numel() * 6, dtype=torch.float64).view(t.numel(), 2, 3) / 10

    y = tdq.odeint(f, y0, t, **kw)
    (y * w).sum().backward()
    g_direct = [y0.grad.clone()] + [p.grad.clone() for p in f.parameters()]
    y0.grad = None
    f.zero_grad()

    ya = tdq.odeint_adjoint(f, y0, t, **kw)
    assert torch.allclose(ya, y.detach(), atol=1e-12)
    (ya * w).sum().backward()
    g_adj = [y0.grad.clone()] + [p.grad.clone() for p in f.parameters()]
    tol = 2e-4 if method == "rk4" else 1e-6      # rk4 on 8 coarse steps: continuous vs discrete adjoint differ at O(h^4)
    for a, b in zip(g_direct, g_adj):
        assert torch.allclose(a, b, rtol=tol, atol=tol), float((a - b).abs().max())


def test_adjoint_seminorm_ignores_parameter_adjoints_in_the_step_control():
    """adjoint.py `handle_adjoint_norm_`: adjoint_options = {"norm": "seminorm"} keeps the parameter adjoints out of the accepted-error
    norm.  The gradients still agree with autograd through the solver (to the tolerance of the solve), and the backward solve takes
    no more attempts than under the default mixed norm -- fewer when the parameter adjoints are what the mixed norm was resolving."""
    f = _Lin()
    calls = {"n": 0}
    fwd = f.forward

    def counted(t, y):
        calls["n"] += 1
        return fwd(t, y)
    f.forward = counted
    y0 = torch.tensor([[0.3, -0.2, 0.5], [0.1, 0.4, -0.6]], dtype=torch.float64, requires_grad=True)
    t = torch.linspace(0, 1.0, 4, dtype=torch.float64)
    w = torch.arange(1, 1 + t.numel() * 6, dtype=torch.float64).view(t.numel(), 2, 3) * 10.0
    kw = dict(method="dopri5", rtol=1e-6, atol=1e-8)
    y = tdq.odeint(f, y0, t, **kw)
    (y * w).sum().backward()
    g_direct = [y0.grad.clone()] + [p.grad.clone() for p in f.parameters()]
    res = {}
    for name, ao in (("mixed", None), ("seminorm", {"norm": "seminorm"})):
        y0.grad = None
        f.zero_grad()
        ya = tdq.odeint_adjoint(f, y0, t, adjoint_options=ao, **kw)
        n0 = calls["n"]
        (ya * w).sum().backward()
        res[name] = (calls["n"] - n0, [y0.grad.clone()] + [p.grad.clone() for p in f.parameters()])
    assert res["seminorm"][0] <= res["mixed"][0], (res["seminorm"][0], res["mixed"][0])
    for name in res:
        for a, b in zip(g_direct, res[name][1]):
            assert torch.allclose(a, b, rtol=2e-4, atol=2e-4 * float(a.abs().max())), (name, float((a - b).abs().max()))


def test_rk4_step_size_grid_and_linear_interpolation():
    """options['step_size'] (solvers.py FixedGridODESolver): the solver steps over t[0] + k h with the last point moved onto t[-1];
    a requested time on a grid point returns the grid row, any other one the LINEAR interpolant of its two neighbours."""
    f = lambda t, y: -y * torch.cos(t)      # noqa: E731
    y0 = torch.tensor([1.0, 2.0], dtype=torch.float64)
    t = torch.linspace(0.0, 2.0, 5, dtype=torch.float64)
    same = tdq.odeint(f, y0, t, method="rk4", options={"step_size": 0.5})
    assert torch.equal(same, tdq.odeint(f, y0, t, method="rk4"))           # grid == t: identical arithmetic
    exact = y0 * torch.exp(-torch.sin(t))[:, None]
    errs = [float((tdq.odeint(f, y0, t, method="rk4", options={"step_size": h}) - exact).abs().max()) for h in (0.25, 0.125)]
    assert 12.0 < errs[0] / errs[1] < 20.0                                  # fourth order in the step size, not in the output spacing
    grid = tdq._grid_from_step_size(torch.tensor([0.0, 2.0], dtype=torch.float64), 0.75)
    assert grid.tolist() == [0.0, 0.75, 1.5, 2.0]                           # shortened last step
    rows = tdq.odeint(f, y0, grid, method="rk4")                            # the same steps, every grid row kept
    tq = torch.tensor([0.0, 0.3, 0.75, 1.8, 2.0], dtype=torch.float64)
    out = tdq.odeint(f, y0, tq, method="rk4", options={"step_size": 0.75})
    assert torch.equal(out[2], rows[1]) and torch.equal(out[4], rows[3])
    assert torch.allclose(out[1], rows[0] + (0.3 / 0.75) * (rows[1] - rows[0]), atol=1e-15)
    assert torch.allclose(out[3], rows[2] + ((1.8 - 1.5) / 0.5) * (rows[3] - rows[2]), atol=1e-15)


def test_adjoint_with_step_size_converges_to_the_exact_gradient():
    """odeint_adjoint(method='rk4', options={'step_size': h}): forward and augmented backward solve on the step_size grid; the
    gradient of y(T) w.r.t. y0 of dy/dt = -y cos t is exp(-sin T), reached at fourth order"""
    class F(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.k = torch.nn.Parameter(torch.tensor(1.0, dtype=torch.float64))

        def forward(self, t, y):
            return -self.k * y * torch.cos(t)
    f = F()
    t = torch.tensor([0.0, 2.0], dtype=torch.float64)
    errs = []
    for h in (0.25, 0.125):
        y0 = torch.tensor([1.5], dtype=torch.float64, requires_grad=True)
        f.zero_grad()
        tdq.odeint_adjoint(f, y0, t, method="rk4", options={"step_size": h})[-1].sum().backward()
        exact_y0 = float(torch.exp(-torch.sin(t[1])))
        exact_k = float(-torch.sin(t[1]) * 1.5 * torch.exp(-torch.sin(t[1])))
        errs.append(max(abs(float(y0.grad) - exact_y0), abs(float(f.k.grad) - exact_k)))
    assert errs[0] < 1e-4 and 10.0 < errs[0] / errs[1] < 40.0      # fourth order (measured ratio 25)
