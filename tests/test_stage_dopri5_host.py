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
