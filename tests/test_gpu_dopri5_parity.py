"""GPU parity of the tensor-core dopri5 path AT THE BENCHMARKED TOLERANCE (rtol = atol = 1e-5, mode_sep/config.py:27-28)
against the CPU oracle (oracle/torchdiffeq_oracle.py, fp32), through the C ABI.

The forward stage kernel of this path (csrc/stage_fwd2_tc.cu, operand format "fp16x2") feeds every activation to the
tensor core as a two-term fp16 split, so the only rounding in a drift evaluation is the fixed fp16 rounding of the
weights.  Consequences tested here:
  * with weights that ARE fp16 numbers the evaluation is fp32-class (<= 5e-6) and the adaptive solver takes EXACTLY the
    oracle's accepted / rejected step sequence;
  * with arbitrary fp32 weights the step count stays within +-15 % of the oracle's (the single-term fp16 / bf16 formats
    take 2.4x / 14x the steps) and the trajectory / gradients stay inside the stated tensor-core tolerance.
"""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import models_oracle as mo
from oracle import torchdiffeq_oracle as tdq

TOL_TRAJ = 5e-3          # stated tolerance of the tensor-core path, relative to the trajectory's max magnitude
# Gradients go through the bf16-operand backward kernels: measured on B200 (profiles/r02_pytest_gpu.log) over 30 accepted steps
# of 6 stages each: dL/dy0 2.0e-2 rms, dL/dW 4.4e-3 rms (worst parameter block); trajectory 4.0e-4 (4.1e-6 with fp16 weights).
TOL_GRAD_RMS = 3e-2


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


def _pair(Z=8, seed=0, fp16_weights=False):
    import ananke_abm_b200 as ab
    torch.manual_seed(seed)
    oracle = mo.OracleModeSep(Z)
    if fp16_weights:      # make every drift-net MATRIX an fp16 number (biases / time columns are applied in fp32 by the kernel)
        with torch.no_grad():
            for m in oracle.odefunc.func.net.modules():
                if isinstance(m, torch.nn.Linear):
                    w = m.weight
                    if w.shape[1] == 162:
                        w[:, :160] = w[:, :160].half().float()
                    else:
                        w.copy_(w.half().float())
    model = ab.ModeSepModel(Z, ab.ModeSepConfig())
    model.load_state_dict(oracle.state_dict())
    return oracle, model


def _agents(B, Z, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, Z, (B,), generator=g), torch.randint(0, Z, (B,), generator=g), torch.rand(B, 2, generator=g)


def _accepted(log):
    return sum(1 for (_, _, ok) in log if ok)


@pytest.mark.parametrize("B", [1, 129, 700])
@pytest.mark.parametrize("fp16_weights", [True, False])
def test_split_forward_single_eval_vs_oracle(B, fp16_weights):
    """one drift evaluation through operand format 2 == WrappedSDE.forward (mode_sep/architecture/model.py:56-73)"""
    import ananke_abm_b200 as ab
    from ananke_abm_b200 import stage
    dev = _cuda()
    oracle, model = _pair(fp16_weights=fp16_weights)
    model = model.to(dev)
    home, work, traits = _agents(B, 8)
    with torch.no_grad():
        y0 = oracle.initial_state(home, work, traits)
        y0[:, 64:128] = 0.2 * torch.randn(B, 64, generator=torch.Generator().manual_seed(3))
        ref = oracle.rhs(torch.tensor(5.25), y0)[:, 64:128]
    spec = ab.describe_drift(model.odefunc)
    eng = stage.TcEngine(spec, spec.flat_params().detach())
    ob = stage.blocked_zeros(B, 64, dev)
    eng.stage_forward(stage.rows_block(y0.to(dev)), [], stage.Combo(0.0, [], []), 5.25, B, a_out=ob, fp16="fp16x2")
    out = stage.rows_unblock(ob, B, 64)
    torch.cuda.synchronize()
    eng.check_status()
    assert not torch.isnan(out).any()
    err = _rel(out.cpu(), ref)
    # fp16 weights: every product is exact in the fp32 accumulator, the evaluation is fp32-class
    assert err < (2e-5 if fp16_weights else 2e-3), err


def test_split_forward_has_no_activation_noise():
    """Evaluations at two nearby states differ by what the fp32 net says they differ by: the DIFFERENCE of two drift
    evaluations (what an embedded error estimate is made of) is reproduced to ~1e-3 of itself, where single-term fp16
    activations lose it in rounding noise."""
    import ananke_abm_b200 as ab
    from ananke_abm_b200 import stage
    dev = _cuda()
    oracle, model = _pair(fp16_weights=True)
    model = model.to(dev)
    B = 512
    home, work, traits = _agents(B, 8)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        y0 = oracle.initial_state(home, work, traits)
        y0[:, 64:128] = 0.2 * torch.randn(B, 64, generator=g)
        y1 = y0.clone()
        y1[:, :128] += 1e-3 * torch.randn(B, 128, generator=g)
        ref = (oracle.rhs(torch.tensor(5.25), y1) - oracle.rhs(torch.tensor(5.25), y0))[:, 64:128]
    spec = ab.describe_drift(model.odefunc)
    eng = stage.TcEngine(spec, spec.flat_params().detach())

    def diff(fmt):
        o = []
        for y in (y0, y1):
            ob = stage.blocked_zeros(B, 64, dev)
            eng.stage_forward(stage.rows_block(y.to(dev)), [], stage.Combo(0.0, [], []), 5.25, B, a_out=ob, fp16=fmt)
            o.append(stage.rows_unblock(ob, B, 64).cpu())
        return o[1] - o[0]
    e_split, e_fp16 = _rms(diff("fp16x2"), ref), _rms(diff("fp16"), ref)
    print(f"difference of two evaluations, rms error relative to the difference: fp16x2 {e_split:.2e}, fp16 {e_fp16:.2e}")
    assert e_split < 2e-2, e_split
    assert e_split < 0.2 * e_fp16, (e_split, e_fp16)


def _solve_both(oracle, model, dev, B, t, tol, opts=None):
    import ananke_abm_b200 as ab
    oi = importlib.import_module("ananke_abm_b200.odeint")
    T = t.numel()
    home, work, traits = _agents(B, 8)
    wgt = torch.linspace(0.5, 1.5, T)[:, None, None]
    oracle.zero_grad()
    y0r = oracle.initial_state(home, work, traits).detach().requires_grad_(True)
    ref = tdq.odeint(oracle.rhs, y0r, t, method="dopri5", rtol=tol, atol=tol)
    log = list(tdq._LAST_SOLVER["solver"].step_log)
    ((ref[:, :, :128] * wgt) ** 2).mean().backward()
    for p in model.parameters():
        p.grad = None
    y0 = model.initial_state(home.to(dev), work.to(dev), traits.to(dev)).detach().requires_grad_(True)
    o = {"precision": "bf16"}
    o.update(opts or {})
    out = ab.odeint(model.odefunc, y0, t.to(dev), method="dopri5", rtol=tol, atol=tol, options=o)
    stats = oi._LAST["solver"]
    ((out[:, :, :128] * wgt.to(dev)) ** 2).mean().backward()
    torch.cuda.synchronize()
    return ref.detach(), y0r.grad, log, out.detach().cpu(), y0.grad.cpu(), stats


def test_dopri5_tc_takes_the_oracle_step_sequence_with_fp16_weights():
    """rtol = atol = 1e-5, B = 2,048: when the weights are fp16 numbers the split-activation forward is an fp32-class
    evaluation of the SAME net, so the controller accepts and rejects exactly as the fp32 oracle does and the dense
    output agrees to round-off amplified by the dynamics."""
    dev = _cuda()
    oracle, model = _pair(fp16_weights=True)
    model = model.to(dev)
    t = torch.linspace(0.0, 24.0, 13)
    ref, gref, log, out, gy0, stats = _solve_both(oracle, model, dev, 2048, t, 1e-5)
    n_acc, n_rej = _accepted(log), len(log) - _accepted(log)
    print(f"fp16-weight net: oracle {n_acc}+{n_rej} steps, tensor-core {stats.n_accepted}+{stats.n_rejected}; "
          f"traj {_rel(out, ref):.2e}, dL/dy0 rms {_rms(gy0, gref):.2e}")
    assert abs(stats.n_accepted - n_acc) <= 1 and abs(stats.n_rejected - n_rej) <= 1      # measured: identical (30 + 4)
    assert torch.equal(out[0], ref[0])
    assert _rel(out, ref) < 1e-4, _rel(out, ref)
    assert _rms(gy0, gref) < TOL_GRAD_RMS


def test_dopri5_tc_parity_at_the_benchmarked_tolerance():
    """rtol = atol = 1e-5, B = 2,048, arbitrary fp32 weights (forward + discrete adjoint): accepted steps within +-15 % of
    the fp32 oracle's, trajectory and gradients inside the stated tensor-core tolerance."""
    dev = _cuda()
    oracle, model = _pair()
    model = model.to(dev)
    t = torch.linspace(0.0, 24.0, 13)
    ref, gref, log, out, gy0, stats = _solve_both(oracle, model, dev, 2048, t, 1e-5)
    n_acc, n_rej = _accepted(log), len(log) - _accepted(log)
    e_traj, e_g = _rel(out, ref), _rms(gy0, gref)
    gw = [(n, _rms(p.grad.cpu(), q.grad)) for (n, p), (_, q) in zip(model.odefunc.func.net.named_parameters(),
                                                                  oracle.odefunc.func.net.named_parameters())]
    print(f"oracle {n_acc}+{n_rej} steps, tensor-core {stats.n_accepted}+{stats.n_rejected}; traj {e_traj:.2e}, "
          f"dL/dy0 rms {e_g:.2e}, dL/dW rms max {max(v for _, v in gw):.2e}")
    assert abs(stats.n_accepted - n_acc) <= max(1, round(0.15 * n_acc)), (stats.n_accepted, n_acc)
    assert stats.n_rejected <= n_rej + max(2, round(0.15 * n_acc))
    assert e_traj < TOL_TRAJ, e_traj
    assert e_g < TOL_GRAD_RMS, e_g
    for n, v in gw:
        assert v < 2e-2, (n, v)


def test_dopri5_single_term_formats_are_noise_limited():
    """the reason operand format 2 exists: at 1e-5 the single-term fp16 forward needs far more steps for the same day"""
    dev = _cuda()
    oracle, model = _pair()
    model = model.to(dev)
    t = torch.linspace(0.0, 24.0, 13)
    _, _, log, _, _, s2 = _solve_both(oracle, model, dev, 512, t, 1e-5)
    _, _, _, _, _, s1 = _solve_both(oracle, model, dev, 512, t, 1e-5, {"forward_operands": "fp16"})
    print(f"accepted steps at 1e-5: oracle {_accepted(log)}, fp16x2 {s2.n_accepted}, fp16 {s1.n_accepted}")
    assert s1.n_accepted > 1.3 * s2.n_accepted
