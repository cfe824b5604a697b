"""SDE branch on the GPU (ab200_sde_euler_step, ananke_abm_b200.sdeint) against oracle/sde_oracle.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import models_oracle as mo
from oracle import sde_oracle as so


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def test_euler_step_kernel_noise_and_update_match_the_specification():
    import ctypes as C  # noqa: F401
    from ananke_abm_b200 import _lib
    dev = _cuda()
    L = _lib.lib()
    B, D = 333, 64
    g = torch.Generator().manual_seed(0)
    y, f = torch.randn(B, D, generator=g).to(dev), torch.randn(B, D, generator=g).to(dev)
    diff = torch.zeros(D, device=dev)
    diff[:32] = 0.1
    out, xi = torch.empty_like(y), torch.empty_like(y)
    rc = L.ab200_sde_euler_step(y.data_ptr(), f.data_ptr(), diff.data_ptr(), 0, B, D, 0.01, 12345678901234, 7, out.data_ptr(),
                                xi.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ab200_sde_euler_step")
    ref_xi = torch.from_numpy(so.normals(B, D, 12345678901234, 7))
    assert float((xi.cpu() - ref_xi).abs().max()) < 2e-5            # same Philox bits; logf / sincosf differ in the last ulps
    ref = y.cpu() + f.cpu() * 0.01 + diff.cpu() * float(np.sqrt(np.float32(0.01))) * ref_xi
    assert float((out.cpu() - ref).abs().max()) < 1e-5
    assert torch.equal(out[:, 32:], (y + f * 0.01)[:, 32:]) or float((out[:, 32:] - (y + f * 0.01)[:, 32:]).abs().max()) < 1e-6


@pytest.mark.parametrize("B", [2, 130])
def test_sdeint_mode_sep_drift_matches_oracle_euler_maruyama(B):
    """the reference's call shape: ScaledSDE(WrappedSDE) through sdeint(method='euler', dt) with unaligned output times"""
    import ananke_abm_b200 as ab
    from ananke_abm_b200.mode_sep import _ScaledSDE
    dev = _cuda()
    torch.manual_seed(42)
    cfg = ab.ModeSepConfig()
    m = ab.ModeSepModel(8, cfg).to(dev)
    om = mo.OracleModeSep(8)
    om.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
    y0 = torch.randn(B, 160, generator=torch.Generator().manual_seed(1)) * 0.3
    ts = torch.tensor([0.0, 0.05, 0.123, 0.2])

    class OracleSDE:
        def f(self, t, y):
            return om.rhs(t, y)

        def g(self, t, y):
            n = torch.zeros_like(y)
            n[:, :128] = 0.05
            return n

    ref = so.sdeint_euler(OracleSDE(), y0, ts, dt=0.01, seed=99)
    with torch.no_grad():
        out = ab.sdeint(_ScaledSDE(m.odefunc, 0.05), y0.to(dev), ts.to(dev), method="euler", dt=0.01, seed=99)
    assert out.shape == ref.shape and torch.equal(out[0].cpu(), y0)
    assert float((out.cpu() - ref).abs().max()) < 2e-5 * float(ref.abs().max())
    assert torch.equal(out[:, :, 128:].cpu(), y0[:, 128:].expand(4, B, 32))      # the context h carries no noise and no drift
    # counter-based noise: agent 0 alone gets the trajectory it had inside the batch
    with torch.no_grad():
        solo = ab.sdeint(_ScaledSDE(m.odefunc, 0.05), y0[:1].to(dev), ts.to(dev), method="euler", dt=0.01, seed=99)
    assert float((solo[:, 0] - out[:, 0]).abs().max()) < 1e-6
    # training through the sampler (the reference's default latent_ode path, latent_ode/train/train.py:57-74): same values as
    # the no-grad call, gradients == autograd through the oracle's Euler-Maruyama loop with the same noise
    for p in m.parameters():
        p.grad = None
    yg = y0.to(dev).requires_grad_(True)
    og = ab.sdeint(_ScaledSDE(m.odefunc, 0.05), yg, ts.to(dev), method="euler", dt=0.01, seed=99)
    assert float((og.detach() - out).abs().max()) < 1e-6
    og[:, :, :128].square().mean().backward()
    om.zero_grad()
    yr = y0.clone().requires_grad_(True)
    rg = so.sdeint_euler(OracleSDE(), yr, ts, dt=0.01, seed=99, with_grad=True)
    rg[:, :, :128].square().mean().backward()
    assert float((yg.grad.cpu() - yr.grad).abs().max()) < 2e-5 * float(yr.grad.abs().max())
    for (n_, p), (_, q) in zip(m.odefunc.func.net.named_parameters(), om.odefunc.func.net.named_parameters()):
        assert float((p.grad.cpu() - q.grad).abs().max()) < 5e-5 * float(q.grad.abs().max()) + 1e-9, n_


def test_mode_sep_model_sde_config_switch():
    import ananke_abm_b200 as ab
    dev = _cuda()
    torch.manual_seed(0)
    cfg = ab.ModeSepConfig()
    cfg.enable_sde, cfg.sde_noise_strength, cfg.sde_dt, cfg.sde_seed = True, 0.01, 0.01, 3
    m = ab.ModeSepModel(8, cfg).to(dev)
    home, work = torch.tensor([1, 2], device=dev), torch.tensor([3, 4], device=dev)
    traits = torch.rand(2, 2, device=dev)
    with torch.no_grad():
        pred_emb, logits, v_t = m(torch.linspace(0, 0.2, 5, device=dev), home, work, traits)
    assert pred_emb.shape == (2, 5, 64) and logits.shape == (2, 5, 8) and torch.isfinite(logits).all()
