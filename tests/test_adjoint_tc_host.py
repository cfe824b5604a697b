"""Host logic of the tensor-core continuous adjoint (adjoint_tc.py) on the CPU: the algorithm runs against a stand-in engine that
evaluates the drift and its vector-Jacobian product with plain torch float64 on the same tile-blocked buffers, and must reproduce
`oracle.odeint_adjoint(method='rk4')` (restated torchdiffeq: adjoint.py + fixed_grid.py) to round-off -- with and without
options['step_size'].  The GPU tests then only have to vouch for the kernels, not for the stage algebra."""
import numpy as np
import pytest
import torch

from ananke_abm_b200 import adjoint_tc, stage
from oracle import models_oracle as mo
from oracle import torchdiffeq_oracle as tdq

TM = stage.TM


class FakeLayout:
    @staticmethod
    def block(src, dst=None, accumulate=False):
        B, F = src.shape
        Bp = stage.padded_rows(B)
        full = torch.zeros(Bp, F, dtype=src.dtype)
        full[:B] = src
        blk = full.view(Bp // TM, TM, F // 4, 4).permute(0, 2, 1, 3).reshape(-1)
        if dst is None:
            return blk.clone()
        if accumulate:
            dst.add_(blk)
        else:
            dst.copy_(blk)
        return dst

    @staticmethod
    def unblock(src, B, F, out=None):
        rows = src.view(-1, F // 4, TM, 4).permute(0, 2, 1, 3).reshape(-1, F)[:B]
        if out is None:
            return rows.clone()
        out.copy_(rows)
        return out

    @staticmethod
    def zeros(B, F, device):
        return torch.zeros(stage.padded_rows(B) * F, dtype=torch.float64)


class FakeEngine:
    """the TcEngine calls adjoint_tc.py makes, in float64 torch"""

    def __init__(self, func, P, H):
        self.func, self.P, self.H, self.D = func, P, H, 2 * P + H
        self.params = tuple(func.parameters())
        self.n_vjp = 0

    def _stage_input(self, y0, a, cin, B):
        P = self.P
        y = FakeLayout.unblock(y0, B, self.D)
        aj = [FakeLayout.unblock(x, B, P) for x in a]
        p0, v0, h = y[:, :P], y[:, P:2 * P], y[:, 2 * P:]
        p = p0 + cin.cpv * v0 + sum(cin.cpa[j] * aj[j] for j in range(len(aj)))
        v = v0 + sum(cin.cva[j] * aj[j] for j in range(len(aj)))
        return torch.cat([p, v, h], dim=1), (p0, v0, h, aj)

    def stage_forward(self, y0, a, cin, t, B, a_out=None, y_out=None, cout=None, **kw):
        P = self.P
        x, (p0, v0, h, aj) = self._stage_input(y0, a, cin, B)
        with torch.no_grad():
            acc = self.func(torch.tensor(t, dtype=torch.float64), x)[:, P:2 * P]
        if a_out is not None:
            FakeLayout.block(acc, a_out)
        if y_out is not None:
            n = len(aj)
            p1 = p0 + cout.cpv * v0 + sum(cout.cpa[j] * aj[j] for j in range(n)) + cout.cpa[n] * acc
            v1 = v0 + sum(cout.cva[j] * aj[j] for j in range(n)) + cout.cva[n] * acc
            FakeLayout.block(torch.cat([p1, v1, h], dim=1), y_out)

    def stage_forward_fused(self, y0, a_bufs, stages, B, y_out=None, cout=None, **kw):
        for k, (n_a, cin, t, a_out) in enumerate(stages):
            last = k == len(stages) - 1
            self.stage_forward(y0, a_bufs[:n_a], cin, t, B, a_out=a_out, y_out=y_out if last else None, cout=cout if last else None)

    def backward_begin(self, B, stages_per_flush):
        self.gw = [torch.zeros_like(p) for p in self.params]

    def stage_backward(self, y0, a, cin, t, B, g_base, gx, dp, dv, gx_out):
        P = self.P
        if len(gx):      # upstream = g_base + sum_l dp[l] gx[l].p + dv[l] gx[l].v
            g_base = g_base.clone()
            uv = g_base.view(-1, P // 4, TM, 4)
            for g, a_, b_ in zip(gx, dp, dv):
                sp_, sv_, _ = adjoint_tc._views(g, self.D, P)
                uv.add_(a_ * sp_ + b_ * sv_)
        x, _ = self._stage_input(y0, a, cin, B)
        x = x.detach().requires_grad_(True)
        with torch.enable_grad():
            acc = self.func(torch.tensor(t, dtype=torch.float64), x)[:, P:2 * P]
            g = torch.autograd.grad(acc, (x,) + self.params, FakeLayout.unblock(g_base, B, P), allow_unused=True)
        FakeLayout.block(g[0], gx_out)
        for acc_, gi in zip(self.gw, g[1:]):
            if gi is not None:
                acc_.add_(gi)
        self.n_vjp += 1

    def combine_backward(self, g, c, B, G_y0, G_a, accumulate):
        """ab200_pv_combine_backward: G_y0 = [g.p, cpv g.p + g.v, g.h] ; G_a[j] = cpa[j] g.p + cva[j] g.v"""
        assert not accumulate
        D, P = self.D, self.P
        gp, gv, gh = adjoint_tc._views(g, D, P)
        op, ov, oh = adjoint_tc._views(G_y0, D, P)
        op.copy_(gp)
        ov.copy_(c.cpv * gp + gv)
        oh.copy_(gh)
        for j, buf in enumerate(G_a):
            buf.view(-1, P // 4, TM, 4).copy_(c.cpa[j] * gp + c.cva[j] * gv)

    def stage_backward_fused(self, y0, a_bufs, stages, B, x_blobs=None, save_level=0, y0_accum=None, upstream=None):
        """ab200_stage_backward_fused without a gather entry: entry k's upstream = g_base + sum over earlier entries' gx"""
        assert y0_accum is None and upstream is None and x_blobs is None
        D, P = self.D, self.P
        for (n_a, cin, t, g_base, sources, gx_out) in stages:
            u = g_base.clone()
            uv = u.view(-1, P // 4, TM, 4)
            for (src, dp, dv) in sources:
                sp_, sv_, _ = adjoint_tc._views(stages[src][5], D, P)
                uv.add_(dp * sp_ + dv * sv_)
            self.stage_backward(y0, a_bufs[:n_a], cin, t, B, u, [], [], [], gx_out)

    def adjoint_gather(self, base, gx, cpv, B, out):
        """ab200_adjoint_gather: out = base + sum_l [gx_l.p, cpv_l gx_l.p + gx_l.v, gx_l.h]"""
        D, P = self.D, self.P
        out.copy_(base)
        op, ov, oh = adjoint_tc._views(out, D, P)
        for g, cv in zip(gx, cpv):
            gp, gv, gh = adjoint_tc._views(g, D, P)
            op.add_(gp)
            ov.add_(cv * gp + gv)
            oh.add_(gh)

    def flush(self):
        pass

    def backward_end(self):
        return torch.cat([g.reshape(-1) for g in self.gw])

    def check_status(self):
        pass


def _setup(B, seed=0):
    torch.manual_seed(seed)
    oracle = mo.OracleModeSep(8).double()
    g = torch.Generator().manual_seed(seed + 1)
    home, work = torch.randint(0, 8, (B,), generator=g), torch.randint(0, 8, (B,), generator=g)
    traits = torch.rand(B, 2, generator=g, dtype=torch.float64)
    y0 = oracle.initial_state(home, work, traits).detach()
    return oracle, y0


def test_step_grid_mirrors_the_package_grid_constructor():
    assert adjoint_tc.step_grid(0.0, 1.0, None) == [0.0, 1.0]
    g = adjoint_tc.step_grid(0.0, 1.0, 0.25, np.float64)
    assert g == [0.0, 0.25, 0.5, 0.75, 1.0]
    g = adjoint_tc.step_grid(1.0, 0.0, 0.25, np.float64)
    assert g == [1.0, 0.75, 0.5, 0.25, 0.0]
    g = adjoint_tc.step_grid(0.0, 1.0, 0.3, np.float64)      # last step is shorter: 0, .3, .6, .9 -> 1.0
    assert len(g) == 5 and g[-1] == 1.0 and abs(g[3] - 0.9) < 1e-12
    ref = tdq._grid_from_step_size(torch.tensor([2.0, 5.5], dtype=torch.float64), 0.4)
    assert np.allclose(adjoint_tc.step_grid(2.0, 5.5, 0.4, np.float64), ref.numpy(), atol=1e-14)
    with pytest.raises(ValueError):
        adjoint_tc.step_grid(0.0, 1.0, 0.0)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("B,step_size", [(3, None), (130, None), (5, 0.25), (5, 0.4)])
def test_tc_continuous_adjoint_algebra_matches_the_oracle(B, step_size, fused):
    oracle, y0 = _setup(B)
    func = oracle.odefunc
    E = oracle.dims.emb_dim if hasattr(oracle, "dims") else 64
    P, H = 64, 32
    t = torch.tensor([0.0, 0.7, 1.0, 2.0], dtype=torch.float64)
    wgt = torch.linspace(0.5, 1.5, t.numel(), dtype=torch.float64)[:, None, None]
    opts = {} if step_size is None else {"step_size": step_size}

    y0r = y0.clone().requires_grad_(True)
    ref = tdq.odeint_adjoint(func, y0r, t, method="rk4", options=dict(opts))
    ((ref * wgt) ** 2).mean().backward()
    ref_gw = torch.cat([p.grad.reshape(-1) for p in func.parameters()])

    eng = FakeEngine(func, P, H)
    th = [float(x) for x in t]
    rows = adjoint_tc.rk4_forward_rows(eng, y0, th, step_size, lay=FakeLayout, np_dtype=np.float64)
    assert torch.allclose(rows, ref.detach(), atol=1e-12, rtol=1e-12)
    grad_rows = (2.0 * rows * wgt * wgt / rows.numel())
    gy0, gw = adjoint_tc.rk4_continuous_adjoint(eng, th, rows, grad_rows, step_size, lay=FakeLayout, np_dtype=np.float64, fused=fused)
    assert torch.allclose(gy0, y0r.grad, atol=1e-12, rtol=1e-9), float((gy0 - y0r.grad).abs().max())
    assert torch.allclose(gw, ref_gw, atol=1e-12, rtol=1e-9), float((gw - ref_gw).abs().max())
    n_steps = sum(len(adjoint_tc.step_grid(th[i], th[i - 1], step_size, np.float64)) - 1 for i in range(1, len(th)))
    assert eng.n_vjp == 4 * n_steps      # one vector-Jacobian product per Runge-Kutta stage, nothing saved per step
