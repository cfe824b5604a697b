"""Host logic of the stage path (no GPU): tableau algebra, dopri5 dense-output weights, blocked-layout arithmetic."""
import torch

from ananke_abm_b200 import stage


def _direct(tab, p0, v0, a, dt, n):
    kp, y_in = [], []
    for i in range(n):
        p = p0 + dt * sum(b * kp[j] for j, b in enumerate(tab.beta[i]))
        v = v0 + dt * sum(b * a[j] for j, b in enumerate(tab.beta[i]))
        y_in.append((p, v))
        kp.append(v)
    return kp, y_in


def test_stage_inputs_and_solution_match_the_tableaus():
    g = torch.Generator().manual_seed(0)
    p0, v0 = torch.randn(5, dtype=torch.float64, generator=g), torch.randn(5, dtype=torch.float64, generator=g)
    a = [torch.randn(5, dtype=torch.float64, generator=g) for _ in range(7)]
    dt = 0.37
    for tab, n in ((stage.RK38, 4), (stage.DOPRI5, 7)):
        kp, y_in = _direct(tab, p0, v0, a, dt, n)
        for i in range(n):
            c = tab.stage_input(i, dt)
            p = p0 + c.cpv * v0 + sum(c.cpa[j] * a[j] for j in range(i))
            v = v0 + sum(c.cva[j] * a[j] for j in range(i))
            assert torch.allclose(p, y_in[i][0], atol=1e-12) and torch.allclose(v, y_in[i][1], atol=1e-12)
        c = tab.combo(tab.b, dt)
        p1 = p0 + dt * sum(b * kp[j] for j, b in enumerate(tab.b))
        v1 = v0 + dt * sum(b * a[j] for j, b in enumerate(tab.b))
        assert torch.allclose(p0 + c.cpv * v0 + sum(c.cpa[j] * a[j] for j in range(n)), p1, atol=1e-12)
        assert torch.allclose(v0 + sum(c.cva[j] * a[j] for j in range(n)), v1, atol=1e-12)
    # FSAL: the 7th dopri5 stage is evaluated at the 5th-order solution; the error weights sum to zero
    assert stage.DOPRI5.beta[6] == stage.DOPRI5.b[:6] and stage.DOPRI5.b[6] == 0.0
    assert abs(sum(stage.DOPRI5.b_err)) < 1e-15 and abs(sum(stage.DOPRI5.b) - 1.0) < 1e-15


def test_dopri5_dense_output_weights_against_the_oracle_interpolant():
    """W(x) reproduces torchdiffeq's quartic (interp.py _interp_fit/_interp_evaluate) written over k_1..k_7"""
    from oracle import torchdiffeq_oracle as tdq
    g = torch.Generator().manual_seed(1)
    y0 = torch.randn(6, dtype=torch.float64, generator=g)
    k = [torch.randn(6, dtype=torch.float64, generator=g) for _ in range(7)]
    dt = 0.8
    y1 = y0 + dt * sum(c * kk for c, kk in zip(stage._DP_C_SOL, k))
    ymid = y0 + dt * sum(c * kk for c, kk in zip(stage._DP_C_MID, k))
    coeffs = tdq._interp_fit(y0, y1, ymid, k[0], k[6], dt)
    for x in (0.0, 0.13, 0.5, 0.77, 1.0):
        ref = tdq._interp_evaluate(coeffs, torch.tensor(2.0, dtype=torch.float64), torch.tensor(2.0 + dt, dtype=torch.float64),
                                   torch.tensor(2.0 + x * dt, dtype=torch.float64))
        w = stage.dopri5_interp_weights(x)
        got = y0 + dt * sum(wj * kk for wj, kk in zip(w, k))
        assert torch.allclose(got, ref, atol=1e-12), x
    assert max(abs(a - b) for a, b in zip(stage.dopri5_interp_weights(1.0), stage._DP_C_SOL)) < 1e-14


def test_blocked_layout_helpers():
    assert stage.padded_rows(1) == 128 and stage.padded_rows(128) == 128 and stage.padded_rows(129) == 256
    z = stage.blocked_zeros(130, 64, "cpu")
    assert z.numel() == 256 * 64 and float(z.abs().sum()) == 0.0
