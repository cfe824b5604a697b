"""GPU parity of the tensor-core continuous adjoint (adjoint_tc.py: `odeint_adjoint(method='rk4', options={'precision': 'bf16'})`)
against `oracle.odeint_adjoint` (restated torchdiffeq adjoint.py + fixed_grid.py rk4), with and without options['step_size'];
stated tolerance of the tensor-core path.  The stage algebra itself is pinned to round-off on the CPU
(tests/test_adjoint_tc_host.py); these tests vouch for the kernels under it and for the dispatch."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import models_oracle as mo
from oracle import torchdiffeq_oracle as tdq

TOL_TRAJ = 5e-3
TOL_GRAD_MAX = 6e-2
TOL_GRAD_RMS = 2e-2


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _rms(a, b):
    a, b = a.double(), b.double()
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30))


def _pair(Z=8, seed=0):
    import ananke_abm_b200 as ab
    torch.manual_seed(seed)
    oracle = mo.OracleModeSep(Z)
    model = ab.ModeSepModel(Z, ab.ModeSepConfig())
    model.load_state_dict(oracle.state_dict())
    return oracle, model


def _agents(B, Z, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, Z, (B,), generator=g), torch.randint(0, Z, (B,), generator=g), torch.rand(B, 2, generator=g)


@pytest.mark.parametrize("fused", [True, False])       # the two launch structures of adjoint_tc.py (default: fused while it fits)
@pytest.mark.parametrize("B,times,step_size", [
    (130, [0.0, 0.4, 1.0, 1.5, 2.0, 3.0], None),       # one 3/8-rule step per output interval, y re-seeded at every row
    (300, [0.0, 3.0], 0.25),                            # 12 steps forward, 12 augmented steps backward, two rows exist
    (129, [0.0, 0.7, 2.0], 0.3),                        # outputs between grid points (linear interpolation), shortened last step
])
def test_tc_continuous_adjoint_rk4_vs_oracle(B, times, step_size, fused):
    import ananke_abm_b200 as ab
    dev = _cuda()
    oracle, model = _pair()
    model = model.to(dev)
    home, work, traits = _agents(B, 8)
    t = torch.tensor(times)
    T = t.numel()
    wgt = torch.linspace(0.5, 1.5, T)[:, None, None]
    opts = {} if step_size is None else {"step_size": step_size}

    y0r = oracle.initial_state(home, work, traits).detach().requires_grad_(True)
    ref = tdq.odeint_adjoint(oracle.odefunc, y0r, t, method="rk4", options=dict(opts))
    ((ref[:, :, :128] * wgt) ** 2).mean().backward()

    y0 = model.initial_state(home.to(dev), work.to(dev), traits.to(dev)).detach().requires_grad_(True)
    out = ab.odeint_adjoint(model.odefunc, y0, t.to(dev), method="rk4", options=dict(opts, precision="bf16", adjoint_fused=fused))
    assert type(out.grad_fn).__name__.startswith("_ContinuousAdjointRK4TC")      # the tensor-core path, not the fp32 augmented solve
    ((out[:, :, :128] * wgt.to(dev)) ** 2).mean().backward()
    torch.cuda.synchronize()

    assert torch.equal(out[0].detach(), y0.detach())
    assert _rel(out.detach().cpu(), ref.detach()) < TOL_TRAJ
    assert _rel(y0.grad.cpu(), y0r.grad) < TOL_GRAD_MAX and _rms(y0.grad.cpu(), y0r.grad) < TOL_GRAD_RMS, \
        (_rel(y0.grad.cpu(), y0r.grad), _rms(y0.grad.cpu(), y0r.grad))
    for (n, p), (_, q) in zip(model.odefunc.func.net.named_parameters(), oracle.odefunc.func.net.named_parameters()):
        assert _rel(p.grad.cpu(), q.grad) < TOL_GRAD_MAX, (n, _rel(p.grad.cpu(), q.grad))
        assert _rms(p.grad.cpu(), q.grad) < TOL_GRAD_RMS, (n, _rms(p.grad.cpu(), q.grad))


def test_tc_continuous_adjoint_agrees_with_the_fp32_kernel_adjoint_and_is_close_to_the_discrete_one():
    """the same call in strict fp32 (augmented system on the FFMA kernels) and the discrete adjoint of the same grid: three
    routes to the gradient of one solve"""
    import ananke_abm_b200 as ab
    dev = _cuda()
    _, model = _pair()
    model = model.to(dev)
    home, work, traits = _agents(256, 8)
    t = torch.linspace(0.0, 2.0, 9, device=dev)
    grads = {}
    for name, fn, opts in (("tc", ab.odeint_adjoint, {"precision": "bf16"}), ("f32", ab.odeint_adjoint, {"precision": "f32"}),
                           ("discrete", ab.odeint, {"precision": "f32"})):
        y0 = model.initial_state(home.to(dev), work.to(dev), traits.to(dev)).detach().requires_grad_(True)
        model.zero_grad(set_to_none=True)
        out = fn(model.odefunc, y0, t, method="rk4", options=opts)
        (out[:, :, :128] ** 2).mean().backward()
        grads[name] = (y0.grad.clone(), torch.cat([p.grad.reshape(-1) for p in model.odefunc.func.net.parameters()]))
    assert _rms(grads["tc"][0], grads["f32"][0]) < TOL_GRAD_RMS and _rms(grads["tc"][1], grads["f32"][1]) < TOL_GRAD_RMS
    # continuous vs discrete adjoint differ by the discretisation error of the adjoint ODE (O(h^4)), not by round-off
    assert _rms(grads["f32"][0], grads["discrete"][0]) < 5e-3


def test_step_size_option_on_odeint_vs_oracle():
    """torchdiffeq's fixed-grid option on the plain (discrete-gradient) solve, strict fp32: grid t[0] + k step_size, linear
    interpolation of the outputs (solvers.py FixedGridODESolver)"""
    import ananke_abm_b200 as ab
    dev = _cuda()
    oracle, model = _pair()
    model = model.to(dev)
    home, work, traits = _agents(70, 8)
    t = torch.tensor([0.0, 0.33, 1.0, 1.9])
    y0r = oracle.initial_state(home, work, traits).detach().requires_grad_(True)
    ref = tdq.odeint(oracle.odefunc, y0r, t, method="rk4", options={"step_size": 0.25})
    (ref[:, :, :128] ** 2).mean().backward()
    y0 = model.initial_state(home.to(dev), work.to(dev), traits.to(dev)).detach().requires_grad_(True)
    out = ab.odeint(model.odefunc, y0, t.to(dev), method="rk4", options={"step_size": 0.25, "precision": "f32"})
    (out[:, :, :128] ** 2).mean().backward()
    assert _rel(out.detach().cpu(), ref.detach()) < 1e-5
    assert _rel(y0.grad.cpu(), y0r.grad) < 1e-4


def test_tc_continuous_adjoint_memory_does_not_grow_with_the_step_count():
    """the scheme's point: no per-step state is kept -- 96 steps cost the memory of 8"""
    import ananke_abm_b200 as ab
    dev = _cuda()
    _, model = _pair()
    model = model.to(dev)
    B = 40_000
    home, work, traits = _agents(B, 8)
    y0 = model.initial_state(home.to(dev), work.to(dev), traits.to(dev)).detach()
    t = torch.tensor([0.0, 24.0], device=dev)
    peaks, grads = [], []
    for step in (3.0, 0.25):
        y = y0.clone().requires_grad_(True)
        model.zero_grad(set_to_none=True)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats(dev)
        out = ab.odeint_adjoint(model.odefunc, y, t, method="rk4", options={"step_size": step, "precision": "bf16"})
        (out[-1, :, :128] ** 2).mean().backward()
        torch.cuda.synchronize()
        peaks.append(torch.cuda.max_memory_allocated(dev))
        grads.append(y.grad)
        assert torch.isfinite(y.grad).all()
    # 88 more steps: a scheme that kept per-step state would add >= 88 x 40,000 x 640 B = 2.2 GB to the ~1 GB peak
    assert peaks[1] <= peaks[0] * 1.10, peaks
