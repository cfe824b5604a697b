"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import models_oracle as mo
from oracle import torchdiffeq_oracle as tdq


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _load_sd(model, g, prefix):
    sd = {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}
    model.load_state_dict(sd, strict=True)
    return model


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_library_loads_on_gpu_box():
    import ananke_abm_b200 as ab
    assert ab.lib().ab200_abi_version() == 1


def test_drift_eval_matches_reference_rhs(golden_rhs):
    import ananke_abm_b200 as ab
    dev = _cuda()
    # mode_sep drift
    m = _load_sd(ab.ModeSepModel(8, ab.ModeSepConfig()), golden_rhs, "ms_sd_").to(dev)
    y = torch.from_numpy(golden_rhs["ms_y"]).to(dev)
    for i in range(3):
        f = m.odefunc(torch.tensor(float(golden_rhs[f"ms_t{i}"])), y).cpu()
        ref = torch.from_numpy(golden_rhs[f"ms_f{i}"])
        assert _rel(f, ref) < 1e-5, (i, _rel(f, ref))
        assert torch.equal(f[:, 128:], torch.zeros_like(f[:, 128:]))
        assert torch.equal(f[:, :64], y[:, 64:128].cpu())
    # latent drift (tanh residual blocks + potential correction), via the framework's own drift module
    lo = _load_sd(mo.OracleLatentODE(8, 7), golden_rhs, "lo_sd_")
    drift = ab.SecondOrderDrift(16, 32, 128, 2, "tanh", potential=(12, 8, 1.0)).to(dev)
    drift.net.load_state_dict(lo.ode_func.net.state_dict())
    y2 = torch.from_numpy(golden_rhs["lo_y"]).to(dev)
    for i in range(3):
        f = drift(torch.tensor(float(golden_rhs[f"lo_t{i}"])), y2).cpu()
        ref = torch.from_numpy(golden_rhs[f"lo_f{i}"])
        assert _rel(f, ref) < 1e-5, (i, _rel(f, ref))


def test_mode_sep_fixture_forward_matches_golden(golden_mode_sep):
    """BASELINE config 1: reference fixtures, B=2, T=91, Z=8, rk4, seed 42."""
    import ananke_abm_b200 as ab
    dev = _cuda()
    g = golden_mode_sep
    m = _load_sd(ab.ModeSepModel(8, ab.ModeSepConfig()), g, "sd_").to(dev)
    t = torch.from_numpy(g["times_union"]).to(dev)
    home, work, traits = (torch.from_numpy(g[k]).to(dev) for k in ("home_idx", "work_idx", "traits"))
    with torch.no_grad():
        y0 = m.initial_state(home, work, traits)
        y_path = m.integrate(y0, t)
        pred, logits, v_t = m.head(y_path)
    assert y_path.shape == (91, 2, 160)
    assert torch.equal(y_path[0], y0)
    assert _rel(y_path.cpu(), torch.from_numpy(g["y_path"])) < 1e-5
    assert _rel(pred.cpu(), torch.from_numpy(g["pred_emb"])) < 1e-5
    assert _rel(v_t.cpu(), torch.from_numpy(g["v_t"])) < 1e-5
    assert np.array_equal(logits.argmax(-1).cpu().numpy(), g["labels"])
    # h is carried unchanged
    assert torch.equal(y_path[:, :, 128:], y0[None, :, 128:].expand(91, -1, -1))


@pytest.mark.parametrize("B,T", [(1, 2), (63, 5), (64, 3), (65, 4), (300, 25)])
def test_rk4_forward_ragged_sizes_vs_oracle(B, T):
    import ananke_abm_b200 as ab
    dev = _cuda()
    torch.manual_seed(11)
    oracle = mo.OracleModeSep(8)
    m = ab.ModeSepModel(8, ab.ModeSepConfig())
    m.load_state_dict(oracle.state_dict())
    m = m.to(dev)
    g = torch.Generator().manual_seed(B * 1000 + T)
    y0 = torch.randn(B, 160, generator=g) * 0.3
    t = torch.sort(torch.rand(T, generator=g) * 24.0).values
    t = t + torch.arange(T) * 1e-3
    ref = tdq.odeint(oracle.rhs, y0, t, method="rk4")
    with torch.no_grad():
        out = ab.odeint(m.odefunc, y0.to(dev), t.to(dev), method="rk4", rtol=1e-5, atol=1e-5)
    assert out.shape == ref.shape
    assert _rel(out.cpu(), ref.detach()) < 1e-5


def test_rk4_single_time_point_returns_y0():
    import ananke_abm_b200 as ab
    dev = _cuda()
    m = ab.ModeSepModel(8, ab.ModeSepConfig()).to(dev)
    y0 = torch.randn(5, 160, device=dev)
    out = ab.odeint(m.odefunc, y0, torch.tensor([3.0], device=dev), method="rk4")
    assert out.shape == (1, 5, 160) and torch.equal(out[0], y0)


def test_errors_mirror_torchdiffeq():
    import ananke_abm_b200 as ab
    dev = _cuda()
    m = ab.ModeSepModel(8, ab.ModeSepConfig()).to(dev)
    y0 = torch.randn(4, 160, device=dev)
    with pytest.raises(AssertionError):
        ab.odeint(m.odefunc, y0, torch.tensor([0.0, 2.0, 1.0], device=dev), method="rk4")
    with pytest.raises(ValueError):
        ab.odeint(m.odefunc, y0, torch.tensor([0.0, 2.0], device=dev), method="no_such_method")
    with pytest.raises(ab.Ab200Error):
        ab.odeint(m.odefunc, y0.cpu(), torch.tensor([0.0, 2.0]), method="rk4")


def test_rk4_backward_matches_autograd_through_oracle():
    import ananke_abm_b200 as ab
    dev = _cuda()
    torch.manual_seed(5)
    oracle = mo.OracleModeSep(8)
    m = ab.ModeSepModel(8, ab.ModeSepConfig())
    m.load_state_dict(oracle.state_dict())
    m = m.to(dev)
    B, T = 45, 7
    g = torch.Generator().manual_seed(3)
    y0 = (torch.randn(B, 160, generator=g) * 0.3)
    t = torch.linspace(0.0, 6.0, T)
    w = torch.randn(T, B, 160, generator=g)

    y0_ref = y0.clone().requires_grad_(True)
    ref = tdq.odeint(oracle.rhs, y0_ref, t, method="rk4")
    (ref * w).sum().backward()

    y0_dev = y0.to(dev).requires_grad_(True)
    out = ab.odeint(m.odefunc, y0_dev, t.to(dev), method="rk4")
    (out * w.to(dev)).sum().backward()

    assert _rel(out.detach().cpu(), ref.detach()) < 1e-5
    assert _rel(y0_dev.grad.cpu(), y0_ref.grad) < 1e-5
    ref_grads = dict(oracle.odefunc.func.net.named_parameters())
    for name, p in m.odefunc.func.net.named_parameters():
        r = ref_grads[name].grad
        assert p.grad is not None, name
        assert _rel(p.grad.cpu(), r) < 2e-5, (name, _rel(p.grad.cpu(), r))


def test_mode_sep_fixture_training_gradients_match_golden(golden_mode_sep):
    """Full reference training loss (mode_sep/train/train.py:101-159) on the fixtures: every parameter gradient."""
    import ananke_abm_b200 as ab
    from tests.ref_losses import mode_sep_training_loss
    dev = _cuda()
    g = golden_mode_sep
    m = _load_sd(ab.ModeSepModel(8, ab.ModeSepConfig()), g, "sd_").to(dev)
    t = torch.from_numpy(g["times_union"]).to(dev)
    home, work, traits = (torch.from_numpy(g[k]).to(dev) for k in ("home_idx", "work_idx", "traits"))
    pred, logits, v = m(t, home, work, traits)
    loss = mode_sep_training_loss(g, pred, logits, v, m.class_table, dev)
    assert abs(float(loss) - float(g["loss_total"])) < 1e-4 * abs(float(g["loss_total"]))
    loss.backward()
    for name, p in m.named_parameters():
        ref = torch.from_numpy(g["grad_" + name])
        got = p.grad.cpu() if p.grad is not None else torch.zeros_like(ref)
        scale = ref.abs().max().clamp_min(1e-12)
        assert float((got - ref).abs().max() / scale) < 5e-5, (name, float((got - ref).abs().max() / scale))


def test_latent_drift_rk4_forward_and_backward_vs_oracle(golden_rhs):
    import ananke_abm_b200 as ab
    dev = _cuda()
    lo = _load_sd(mo.OracleLatentODE(8, 7), golden_rhs, "lo_sd_")
    drift = ab.SecondOrderDrift(16, 32, 128, 2, "tanh", potential=(12, 8, 1.0)).to(dev)
    drift.net.load_state_dict(lo.ode_func.net.state_dict())
    g = torch.Generator().manual_seed(9)
    B, T = 37, 6
    y0 = torch.randn(B, 64, generator=g) * 0.4
    t = torch.linspace(0, 3.0, T)
    w = torch.randn(T, B, 64, generator=g)
    y0r = y0.clone().requires_grad_(True)
    ref = tdq.odeint(lo.rhs, y0r, t, method="rk4")
    (ref * w).sum().backward()
    y0d = y0.to(dev).requires_grad_(True)
    out = ab.odeint(drift, y0d, t.to(dev), method="rk4")
    (out * w.to(dev)).sum().backward()
    assert _rel(out.detach().cpu(), ref.detach()) < 1e-5
    assert _rel(y0d.grad.cpu(), y0r.grad) < 2e-5
    ref_grads = dict(lo.ode_func.net.named_parameters())
    for name, p in drift.net.named_parameters():
        assert _rel(p.grad.cpu(), ref_grads[name].grad) < 5e-5, name


def test_generic_func_rk4_and_dopri5_vs_oracle():
    import ananke_abm_b200 as ab
    dev = _cuda()

    class F(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(0)
            self.a = torch.nn.Linear(6, 6)

        def forward(self, t, y):
            return torch.tanh(self.a(y)) * torch.cos(t)
    f_cpu = F()
    f_dev = F().to(dev)
    y0 = torch.randn(50, 6, generator=torch.Generator().manual_seed(1))
    t = torch.linspace(0, 2.0, 9)
    ref = tdq.odeint(f_cpu, y0, t, method="rk4")
    out = ab.odeint(f_dev, y0.to(dev), t.to(dev), method="rk4")
    assert _rel(out.detach().cpu(), ref.detach()) < 1e-5
    ref5 = tdq.odeint(f_cpu, y0, t, method="dopri5", rtol=1e-5, atol=1e-6)
    n_ref = tdq._LAST_SOLVER["solver"].n_accepted
    out5 = ab.odeint(f_dev, y0.to(dev), t.to(dev), method="dopri5", rtol=1e-5, atol=1e-6)
    import importlib
    oi = importlib.import_module("ananke_abm_b200.odeint")
    assert _rel(out5.detach().cpu(), ref5.detach()) < 1e-4
    assert abs(oi._LAST["solver"].n_accepted - n_ref) <= 2
    # decreasing time grid
    tr = torch.linspace(2.0, 0.0, 9)
    refr = tdq.odeint(f_cpu, y0, tr, method="rk4")
    outr = ab.odeint(f_dev, y0.to(dev), tr.to(dev), method="rk4")
    assert _rel(outr.detach().cpu(), refr.detach()) < 1e-5


def test_umma_probe_all_operand_layouts():
    """tcgen05 plumbing self-test: every A/B placement and shared-memory layout the kernels may use."""
    import ananke_abm_b200 as ab
    dev = _cuda()
    L = ab.lib()
    torch.manual_seed(0)
    for (N, K) in [(128, 128), (64, 128), (128, 192)]:
        A = torch.randn(128, K, device=dev)
        B = torch.randn(N, K, device=dev)
        ref = A.bfloat16().float() @ B.bfloat16().float().T
        for a_mode in (0, 1, 2):
            for b_mode in (1, 2):
                D = torch.zeros(128, N, device=dev)
                st = torch.full((1,), -7, dtype=torch.int32, device=dev)
                rc = L.ab200_debug_umma_probe(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, a_mode, b_mode, st.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                assert rc == 0 and int(st) == 0, (N, K, a_mode, b_mode, rc, int(st))
                assert float((D - ref).abs().max()) < 1e-3, (N, K, a_mode, b_mode)


# Stated tolerance of the tensor-core (bf16 operands, fp32 accumulate, fp32 state) path against the strict path:
# 2e-2 of the trajectory's scale over a full 96-step day, and >= 99% identical predicted zone labels.
BF16_TRAJ_TOL = 2e-2


@pytest.mark.parametrize("B,T", [(1, 3), (127, 4), (128, 2), (129, 9), (1000, 97)])
def test_rk4_bf16_tensor_core_path_within_stated_tolerance(B, T):
    import ananke_abm_b200 as ab
    dev = _cuda()
    torch.manual_seed(42)
    m = ab.ModeSepModel(50, ab.ModeSepConfig()).to(dev)
    g = torch.Generator().manual_seed(B + T)
    home = torch.randint(0, 50, (B,), generator=g).to(dev)
    work = torch.randint(0, 50, (B,), generator=g).to(dev)
    traits = torch.rand(B, 2, generator=g).to(dev)
    t = torch.linspace(0.0, 24.0 * (T - 1) / 96.0, T, device=dev)
    with torch.no_grad():
        y0 = m.initial_state(home, work, traits)
        ref = ab.odeint(m.odefunc, y0, t, method="rk4", options={"precision": "f32"})
        out = ab.odeint(m.odefunc, y0, t, method="rk4", options={"precision": "bf16"})
    assert out.shape == ref.shape and not torch.isnan(out).any()
    assert torch.equal(out[0], y0)
    assert torch.equal(out[:, :, 128:], ref[:, :, 128:])          # h is carried exactly
    assert _rel(out, ref) < BF16_TRAJ_TOL, _rel(out, ref)
    lab_ref = m.head(ref)[1].argmax(-1)
    lab = m.head(out)[1].argmax(-1)
    assert float((lab == lab_ref).float().mean()) >= 0.99


@pytest.mark.parametrize("n,offset", [(160 * 7, 0), (160 * 7 + 3, 0), (1001, 1), (5, 0), (4096 * 33 + 2, 2)])
def test_combine_errnorm_kernel_bitexact_solution_and_norm(n, offset):
    """ab200_rk_combine_errnorm through the C ABI: y1 is bit-identical to the reference's op order
    (k_j * (c_j * dt) summed stage by stage, then added to y0: torchdiffeq rk_common.py) on the 128-bit path, on the
    scalar tail and on misaligned pointers; the squared error-ratio sum agrees with float64 to summation round-off."""
    import ctypes as C
    from ananke_abm_b200 import _lib
    dev = _cuda()
    L = _lib.lib()
    g = torch.Generator().manual_seed(n)
    pool = [torch.randn(n + 4, generator=g).to(dev) for _ in range(8)]
    y0, ks = pool[0][offset:offset + n], [p[offset:offset + n] for p in pool[1:]]
    csol = [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0]
    cerr = [35 / 384 - 1951 / 21600, 0.0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
            -2187 / 6784 + 12231 / 42400, 11 / 84 - 649 / 6300, -1 / 60]
    dt, rtol, atol = 0.37, 1e-5, 1e-5
    y1 = torch.empty(n + 4, device=dev)[offset:offset + n]
    sumsq = torch.zeros(1, device=dev)
    ptrs = (C.c_void_p * 8)(*([k.data_ptr() for k in ks] + [0]))
    a_sol = (C.c_float * 8)(*(csol + [0.0]))
    a_err = (C.c_float * 8)(*(cerr + [0.0]))
    rc = L.ab200_rk_combine_errnorm(y0.data_ptr(), C.cast(ptrs, C.c_void_p), C.cast(a_sol, C.c_void_p), C.cast(a_err, C.c_void_p), 7,
                                    dt, rtol, atol, y1.data_ptr(), sumsq.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ab200_rk_combine_errnorm")
    f32 = lambda v: torch.tensor(v, dtype=torch.float32, device=dev)   # noqa: E731
    s = torch.zeros(n, device=dev)
    e = torch.zeros(n, device=dev)
    for j in range(7):
        s = s + ks[j] * (f32(csol[j]) * f32(dt))
        e = e + ks[j] * (f32(cerr[j]) * f32(dt))
    ref_y1 = y0 + s
    assert torch.equal(y1, ref_y1)
    tol = atol + rtol * torch.maximum(y0.abs(), ref_y1.abs())
    ref_sum = float(((e / tol).double() ** 2).sum())
    assert abs(float(sumsq[0]) - ref_sum) <= 2e-5 * ref_sum + 1e-12


def test_adjoint_seam_equals_odeint_on_the_stage_path():
    """config.adjoint routes GATODEModel.integrate through odeint_adjoint (configs[4]); with the explicit
    adjoint_mode='discrete' opt-in both seams run the same discrete adjoint, so outputs and gradients are identical."""
    import ananke_abm_b200 as ab
    from ananke_abm_b200.graph import synthetic_zone_graph
    dev = _cuda()
    res = []
    for adjoint in (False, True):
        torch.manual_seed(7)
        mc = ab.ModeSepConfig()
        mc.precision, mc.ode_method, mc.adjoint = "bf16", "dopri5", adjoint
        mc.adjoint_mode = "discrete"
        model = ab.GATODEModel(7, mc, heads=4).to(dev)
        ei, feats = synthetic_zone_graph(64, k=6, seed=1)
        csr = ab.build_zone_csr(ei, 64).to(dev)
        g = torch.Generator().manual_seed(3)
        home, work = torch.randint(0, 64, (300,), generator=g).to(dev), torch.randint(0, 64, (300,), generator=g).to(dev)
        traits = torch.rand(300, 2, generator=g).to(dev)
        table, zemb = model.zone_tables(feats.to(dev), csr)
        y0 = model.initial_state(table, zemb, home, work, traits)
        yp = model.integrate(y0, torch.linspace(0.0, 2.0, 5, device=dev))
        yp[:, :, :128].square().mean().backward()
        res.append((yp.detach().clone(), [p.grad.clone() for p in model.odefunc.parameters()]))
    assert torch.equal(res[0][0], res[1][0])
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-7)     # wgrad partials are summed by atomics: order may differ


@pytest.mark.parametrize("method", ["rk4", "dopri5"])
@pytest.mark.parametrize("B", [0, 1])
def test_empty_and_single_agent_batches(method, B):
    """Edge cases of the batch axis: an empty shard returns [T, 0, D]; one agent equals row 0 of a larger batch."""
    import ananke_abm_b200 as ab
    dev = _cuda()
    torch.manual_seed(5)
    m = ab.ModeSepModel(8, ab.ModeSepConfig()).to(dev)
    t = torch.linspace(0.0, 1.0, 4, device=dev)
    y_big = torch.randn(130, 160, device=dev) * 0.3
    with torch.no_grad():
        out = ab.odeint(m.odefunc, y_big[:B], t, method=method, rtol=1e-5, atol=1e-5)
        assert out.shape == (4, B, 160)
        if B == 1 and method == "rk4":
            ref = ab.odeint(m.odefunc, y_big, t, method=method)
            assert torch.equal(out[:, 0], ref[:, 0])
