"""Host logic of the tensor-core dopri5 driver (`ananke_abm_b200.stage.dopri5_forward`) on the CPU.

The driver's arithmetic lives in the C-ABI kernels (`ab200_dopri5_attempt`, `ab200_dopri5_dense_rows`, `ab200_stage_forward`,
`ab200_rows_block / _unblock`); here they are replaced by a torch stand-in of the SAME contracts over a simple second-order drift,
so that what the HOST decides -- initial step size, accept / reject, step-size control, which requested times fall into which
accepted step, the deferred dense-output launches and the rotation of the (y, a_j) buffer sets at save level 2 -- can be compared
with the oracle solver (oracle/torchdiffeq_oracle.py) without a GPU.  The kernels themselves are tested on the GPU."""
import math

import pytest
import torch

from oracle import torchdiffeq_oracle as tdq

P, H = 4, 2
D = 2 * P + H
W = torch.tensor([[0.3, -1.1, 0.2, 0.0], [0.9, 0.1, -0.4, 0.6], [-0.5, 0.7, 0.2, -0.3], [0.1, 0.2, -0.8, 0.4]])


def _accel(t, p, v, h):
    """a = f(t, p, v, h): smooth, nonlinear, time dependent"""
    return torch.tanh(p @ W.T) * 1.5 - 0.3 * v + 0.2 * math.sin(0.7 * t) + 0.1 * h.sum(-1, keepdim=True)


def _rhs(t, y):
    p, v, h = y[:, :P], y[:, P:2 * P], y[:, 2 * P:]
    return torch.cat([v, _accel(float(t), p, v, h), torch.zeros_like(h)], dim=-1)


class _FakeEngine:
    """torch stand-in of TcEngine's forward entry points; 'blocked' buffers are flat row-major [Bp, F] here"""

    def __init__(self, stage):
        self.st, self.D, self.P, self.fwd_format = stage, D, P, 2
        self.dev = torch.device("cpu")
        self.attempts, self.saved_levels = 0, []

    def _rows(self, buf, F):
        return buf.view(-1, F)

    def _state(self, y0, a, c):
        y = self._rows(y0, D)
        p, v, h = y[:, :P].clone(), y[:, P:2 * P].clone(), y[:, 2 * P:]
        p = p + c.cpv * v
        for j, aj in enumerate(a):
            p = p + c.cpa[j] * self._rows(aj, P)
            v = v + c.cva[j] * self._rows(aj, P)
        return p, v, h

    def stage_forward(self, y0, a, cin, t, B, a_out=None, **kw):
        p, v, h = self._state(y0, a, cin)
        self._rows(a_out, P).copy_(_accel(t, p, v, h))

    def xblob_bytes(self, B, level):
        return 64 * level

    def dopri5_attempt(self, y0, A, t0, dt, B, y_out, err_sumsq, rtol, atol, x_blobs=None, save_level=0):
        st = self.st
        self.attempts += 1
        self.saved_levels.append(save_level)
        if x_blobs is not None:
            x_blobs.fill_(float(self.attempts))      # what an attempt saved is identified by its sequence number
        for i in range(1, 7):                        # stages 2..7 (a_1 is given)
            c = st.DOPRI5.stage_input(i, dt)
            self.stage_forward(y0, A[:i], c, t0 + st.DOPRI5.c[i] * dt, B, a_out=A[i])
        csol = st.DOPRI5.combo(st.DOPRI5.b, dt)
        p, v, h = self._state(y0, A[:7], csol)
        y1 = torch.cat([p, v, h], dim=-1)
        self._rows(y_out, D).copy_(y1)
        # embedded error  dt * sum_j c_err[j] k_j  with k_j = (v_j, a_j): the same linear combination without the base state
        cerr = st.DOPRI5.combo(st._DP_C_ERR, dt)
        y = self._rows(y0, D)
        ep = cerr.cpv * y[:, P:2 * P]
        ev = torch.zeros_like(ep)
        for j in range(7):
            ep = ep + cerr.cpa[j] * self._rows(A[j], P)
            ev = ev + cerr.cva[j] * self._rows(A[j], P)
        tol_p = atol + rtol * torch.maximum(y[:, :P].abs(), y1[:, :P].abs())
        tol_v = atol + rtol * torch.maximum(y[:, P:2 * P].abs(), y1[:, P:2 * P].abs())
        err_sumsq += float(((ep / tol_p).double() ** 2).sum() + ((ev / tol_v).double() ** 2).sum())

    def dopri5_dense_rows(self, y0, A, dt, xs, B, outs):
        st = self.st
        for x, o in zip(xs, outs):
            c = st.DOPRI5.combo(st.dopri5_interp_weights(x), dt)
            p, v, h = self._state(y0, A[:7], c)
            o.copy_(torch.cat([p, v, h], dim=-1)[:B])

    def check_status(self):
        pass


@pytest.fixture()
def fake_stage(monkeypatch):
    from ananke_abm_b200 import stage
    monkeypatch.setattr(stage, "padded_rows", lambda B: B)
    monkeypatch.setattr(stage, "rows_block", lambda src, dst=None, accumulate=False: src.contiguous().reshape(-1).clone())
    monkeypatch.setattr(stage, "rows_unblock", lambda src, B, F, out=None: src.view(-1, F)[:B].clone())
    monkeypatch.setattr(stage, "blocked_zeros", lambda B, F, device: torch.zeros(B * F))
    return stage


@pytest.mark.parametrize("save", [("none", False), ("all", True), ("inputs", True)])
@pytest.mark.parametrize("T", [4, 23])
def test_host_loop_takes_the_oracle_step_sequence_and_dense_output(fake_stage, save, T):
    stage = fake_stage
    mode, save_steps = save
    torch.manual_seed(3)
    B = 5
    y0 = torch.cat([torch.randn(B, P), 0.5 * torch.randn(B, P), torch.randn(B, H)], dim=-1)
    t = torch.linspace(0.0, 6.0, T)
    ref = tdq.odeint(_rhs, y0, t, method="dopri5", rtol=1e-5, atol=1e-6)
    log = tdq._LAST_SOLVER["solver"].step_log
    eng = _FakeEngine(stage)
    y_path, steps, stats = stage.dopri5_forward(eng, y0, [float(x) for x in t], 1e-5, 1e-6, save_steps=save_steps, saved_operands=mode)
    n_acc, n_rej = sum(1 for x in log if x[2]), sum(1 for x in log if not x[2])
    assert (stats.n_accepted, stats.n_rejected) == (n_acc, n_rej)
    assert eng.attempts == n_acc + n_rej
    # every requested time is produced once, from the right step, with the buffers of THAT step (deferred launches + rotation)
    assert float((y_path - ref).abs().max()) < 1e-4 * float(ref.abs().max())      # fp32 round-off in a different association order, amplified by the dynamics
    if save_steps:
        level = {"all": 2, "inputs": 1}[mode]
        assert set(eng.saved_levels) == {level}
        assert len(steps) == n_acc and sorted(k for s in steps for k, _ in s.outputs) == list(range(1, T))
        # the saved buffer of a step is the one its ACCEPTED attempt wrote (rejected attempts are overwritten by the retry)
        accepted_attempt_ids = [i + 1 for i, x in enumerate(log) if x[2]]
        assert [int(s.x[0]) for s in steps] == accepted_attempt_ids
        assert all(s.save_level == level for s in steps)
        # step bookkeeping: consecutive, covering [t0, >= t_end]
        for a, b in zip(steps, steps[1:]):
            assert abs((a.t0 + a.dt) - b.t0) < 1e-12
        assert steps[0].t0 == 0.0 and steps[-1].t0 + steps[-1].dt >= 6.0 - 1e-12
    else:
        assert steps is None and set(eng.saved_levels) == {0}


class _FakeTrainEngine(_FakeEngine):
    """+ the backward entry points (torch stand-ins of the same contracts): `combine_backward_multi`, `stage_backward_fused` with
    gather sources, the GATHER entry (dL/dy0 in place, FSAL hand-over), `flush` / `backward_end`.  The drift has no parameters, so
    the weight gradient is empty; dL/dy0 is what the host algebra produces."""

    def stage_forward_fused(self, y0, a_bufs, stages, B, y_out=None, cout=None, **kw):
        for (n_a, cin, t, a_out) in stages:
            p, v, h = self._state(y0, a_bufs[:n_a], cin)
            acc = _accel(t, p, v, h)
            if a_out is not None:
                self._rows(a_out, P).copy_(acc)
        if y_out is not None:      # the last stage's acceleration closes the step: out coefficients over (a_1..a_n, a_new)
            n = stages[-1][0]
            y = self._rows(y0, D)
            p1 = y[:, :P] + cout.cpv * y[:, P:2 * P] + cout.cpa[n] * acc
            v1 = y[:, P:2 * P] + cout.cva[n] * acc
            for j in range(n):
                p1 = p1 + cout.cpa[j] * self._rows(a_bufs[j], P)
                v1 = v1 + cout.cva[j] * self._rows(a_bufs[j], P)
            self._rows(y_out, D).copy_(torch.cat([p1, v1, y[:, 2 * P:]], dim=-1))

    def backward_begin(self, B, stages_per_flush):
        self.n_vjp = 0

    def combine_backward_multi(self, sources, B, G_y0, G_a, accumulate, add_a=None, add_index=0):
        gy = self._rows(G_y0, D)
        if not accumulate:
            gy.zero_()
            for ga in G_a:
                ga.zero_()
        for g, c in sources:
            g = g if g.dim() == 2 else self._rows(g, D)
            gp, gv, gh = g[:, :P], g[:, P:2 * P], g[:, 2 * P:]
            gy[:, :P] += gp
            gy[:, P:2 * P] += c.cpv * gp + gv
            gy[:, 2 * P:] += gh
            for j, ga in enumerate(G_a):
                self._rows(ga, P).add_(c.cpa[j] * gp + c.cva[j] * gv)
        if add_a is not None:
            G_a[add_index].add_(add_a)

    def stage_backward_fused(self, y0, a_bufs, stages, B, x_blobs=None, save_level=0, y0_accum=None, upstream=None):
        gxs = []
        for (n_a, cin, t, g_base, sources, gx_out) in stages:
            u = torch.zeros(B, P) if g_base is None else self._rows(g_base, P).clone()
            for (src, dp, dv) in sources:
                g = gxs[src]
                u = u + dp * g[:, :P] + dv * g[:, P:2 * P]
            p, v, h = self._state(y0, a_bufs[:n_a], cin)
            x = [z.detach().requires_grad_(True) for z in (p, v, h)]
            with torch.enable_grad():
                acc = _accel(t, *x)
                g = torch.autograd.grad(acc, x, u)
            gx = torch.cat(g, dim=-1)
            self._rows(gx_out, D).copy_(gx)
            gxs.append(gx)
            self.n_vjp += 1
        cpvs = [s[1].cpv for s in stages]
        if y0_accum is not None:      # gather entry: dL/dy0 += sum_s [gx_s.p, cpv_s gx_s.p + gx_s.v, gx_s.h]
            gy = self._rows(y0_accum, D)
            for gx, cpv in zip(gxs, cpvs):
                gy[:, :P] += gx[:, :P]
                gy[:, P:2 * P] += cpv * gx[:, :P] + gx[:, P:2 * P]
                gy[:, 2 * P:] += gx[:, 2 * P:]
        if upstream is not None:      # gradient handed to the previous step's FSAL evaluation
            g_base, coef, out = upstream
            u = self._rows(g_base, P).clone()
            for gx, (dp, dv) in zip(gxs, coef):
                u = u + dp * gx[:, :P] + dv * gx[:, P:2 * P]
            self._rows(out, P).copy_(u)

    def flush(self):
        pass

    def backward_end(self):
        return torch.zeros(1)


@pytest.fixture()
def fake_stage_train(fake_stage, monkeypatch):
    def rows_block(src, dst=None, accumulate=False):
        flat = src.contiguous().reshape(-1)
        if dst is None:
            return flat.clone()
        if accumulate:
            dst.add_(flat)
        else:
            dst.copy_(flat)
        return dst

    def rows_unblock(src, B, F, out=None):
        rows = src.view(-1, F)[:B]
        if out is None:
            return rows.clone()
        out.copy_(rows)
        return out
    monkeypatch.setattr(fake_stage, "rows_block", rows_block)
    monkeypatch.setattr(fake_stage, "rows_unblock", rows_unblock)
    monkeypatch.setattr(fake_stage, "blocked_empty", lambda B, F, device: torch.zeros(B * F))
    return fake_stage


def _weighted_loss(y_path):
    w = torch.linspace(0.5, 1.5, y_path.shape[0])[:, None, None]
    return ((y_path * w) ** 2).mean()


def test_rk4_host_backward_equals_autograd_through_the_oracle(fake_stage_train):
    """stage.rk4_forward / rk4_backward (one elementwise pass over the row-major dL/dy_path row + the later steps' gradient, fused
    backward stages with a gather entry) == reverse-mode autograd through the oracle's rk4 on the same drift"""
    stage = fake_stage_train
    torch.manual_seed(5)
    B, T = 6, 7
    y0 = torch.cat([torch.randn(B, P), 0.5 * torch.randn(B, P), torch.randn(B, H)], dim=-1)
    t = torch.linspace(0.0, 2.0, T)
    y0r = y0.clone().requires_grad_(True)
    ref = tdq.odeint(_rhs, y0r, t, method="rk4")
    _weighted_loss(ref).backward()
    eng = _FakeTrainEngine(stage)
    y_path, saved = stage.rk4_forward(eng, y0, [float(x) for x in t], save_stages=True)
    assert float((y_path - ref.detach()).abs().max()) < 2e-6 * float(ref.detach().abs().max())
    yp = y_path.clone().requires_grad_(True)
    _weighted_loss(yp).backward()
    gy0, _ = stage.rk4_backward(eng, [float(x) for x in t], saved, yp.grad)
    assert float((gy0 - y0r.grad).abs().max()) < 5e-6 * float(y0r.grad.abs().max())
    assert eng.n_vjp == 4 * (T - 1)


@pytest.mark.parametrize("T", [4, 23])
def test_dopri5_host_backward_equals_autograd_through_the_oracle(fake_stage_train, T):
    """stage.dopri5_forward(save_steps) / dopri5_backward: discrete adjoint of the accepted steps with the FSAL evaluation
    differentiated once, dense-output rows read in place == autograd through the oracle's dopri5 (step sizes constant, as there)"""
    stage = fake_stage_train
    torch.manual_seed(3)
    B = 5
    y0 = torch.cat([torch.randn(B, P), 0.5 * torch.randn(B, P), torch.randn(B, H)], dim=-1)
    t = torch.linspace(0.0, 6.0, T)
    # torchdiffeq runs its initial step-size heuristic with autograd on (misc.py `_select_initial_step`), every later step size under
    # no_grad; the drivers here treat dt_0 as a constant too.  Reference = the oracle with the SAME first step given
    # as a number; the term the free-running oracle adds through d(dt_0)/d(y0) is measured below (1.4e-3 of the gradient at T = 4).
    y0f = y0.clone().requires_grad_(True)
    _weighted_loss(tdq.odeint(_rhs, y0f, t, method="dopri5", rtol=1e-5, atol=1e-6)).backward()
    dt0 = float(tdq._LAST_SOLVER["solver"].step_log[0][1])
    y0r = y0.clone().requires_grad_(True)
    ref = tdq.odeint(_rhs, y0r, t, method="dopri5", rtol=1e-5, atol=1e-6, options={"first_step": dt0})
    _weighted_loss(ref).backward()
    eng = _FakeTrainEngine(stage)
    y_path, steps, stats = stage.dopri5_forward(eng, y0, [float(x) for x in t], 1e-5, 1e-6, save_steps=True, saved_operands="none")
    yp = y_path.clone().requires_grad_(True)
    _weighted_loss(yp).backward()
    gy0, _ = stage.dopri5_backward(eng, steps, yp.grad)
    scale = float(y0r.grad.abs().max())
    assert float((gy0 - y0r.grad).abs().max()) < 3e-4 * scale        # fp32 association order, amplified by the dynamics (the oracle's own
    #                                                                   fp32 and fp64 gradients differ by 1.9e-4 of the scale here)
    assert float((y0f.grad - y0r.grad).abs().max()) < 5e-3 * scale    # the initial-step term of the free-running oracle
