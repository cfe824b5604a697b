"""Fused cross-entropy head against the reference formula: cosine logits / tau
(mode_sep/architecture/model.py:196-199) into `ce_at_snaps` (mode_sep/architecture/losses.py:14-22)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _ref_logits(emb, table, tau):
    tn = table / (table.norm(dim=-1, keepdim=True) + 1e-8)
    en = emb / (emb.norm(dim=-1, keepdim=True) + 1e-8)
    return en @ tn.T / tau


@pytest.mark.parametrize("M,Z", [(1, 8), (130, 8), (777, 500), (4096, 10_000), (300, 129)])
def test_head_ce_rows_match_float64_reference(M, Z):
    import ananke_abm_b200 as ab
    dev = _cuda()
    g = torch.Generator().manual_seed(M + Z)
    emb = torch.randn(M, 64, generator=g).to(dev)
    table = torch.randn(Z, 64, generator=g).to(dev)
    tgt = torch.randint(0, Z, (M,), generator=g).to(dev)
    rows = ab.head_ce_rows(emb, table, tgt, 0.2)
    ref = F.cross_entropy(_ref_logits(emb.double(), table.double(), 0.2), tgt, reduction="none")
    err = float((rows.double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err


def test_ce_at_snaps_fused_value_and_gradients_match_reference_formula():
    import ananke_abm_b200 as ab
    dev = _cuda()
    g = torch.Generator().manual_seed(11)
    B, T, Z = 37, 9, 500
    emb = torch.randn(B, T, 64, generator=g).to(dev).requires_grad_(True)
    table = torch.randn(Z, 64, generator=g).to(dev).requires_grad_(True)
    y = torch.randint(0, Z, (B, T), generator=g).to(dev)
    mask = (torch.rand(B, T, generator=g) < 0.3).to(dev)
    y = torch.where(mask, y, torch.full_like(y, -1))          # -1 where not ground truth, as UnionBatch has it
    loss = ab.ce_at_snaps_fused(emb, table, y, mask, 0.2)
    loss.backward()
    e2, t2 = emb.detach().double().requires_grad_(True), table.detach().double().requires_grad_(True)
    logits = _ref_logits(e2, t2, 0.2)
    ref = F.cross_entropy(logits[mask], y[mask], reduction="mean")
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    for a, b in ((emb.grad, e2.grad), (table.grad, t2.grad)):
        assert float((a.double() - b).abs().max()) < 2e-5 * float(b.abs().max())
    # empty mask -> 0, like the reference
    assert float(ab.ce_at_snaps_fused(emb.detach(), table.detach(), y, torch.zeros_like(mask), 0.2)) == 0.0


def test_head_ce_labels_equal_head_argmax():
    """The cross-entropy pass can hand back the argmax labels of the same sweep (C ABI)."""
    import ctypes as C  # noqa: F401
    from ananke_abm_b200 import _lib
    from ananke_abm_b200.inference import head_argmax
    dev = _cuda()
    L = _lib.lib()
    g = torch.Generator().manual_seed(2)
    M, Z = 1000, 700
    emb, table = torch.randn(M, 64, generator=g).to(dev), torch.randn(Z, 64, generator=g).to(dev)
    tgt = torch.randint(0, Z, (M,), generator=g).to(dev)
    ws = torch.empty(int(L.ab200_head_workspace_bytes(Z, 64)), dtype=torch.uint8, device=dev)
    lse, tl = torch.empty(M, device=dev), torch.empty(M, device=dev)
    labels = torch.empty(M, dtype=torch.int64, device=dev)
    rc = L.ab200_head_ce_forward(emb.data_ptr(), table.data_ptr(), tgt.data_ptr(), M, Z, 64, 0.2, lse.data_ptr(), tl.data_ptr(),
                                 labels.data_ptr(), None, None, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ab200_head_ce_forward")
    assert torch.equal(labels, head_argmax(emb, table, 0.2))
    logits = _ref_logits(emb.double(), table.double(), 0.2)
    assert float((tl.double() - logits.gather(1, tgt[:, None])[:, 0]).abs().max()) < 1e-5


@pytest.mark.parametrize("M,Z", [(20_000, 300), (300, 20_000), (5_000, 1_000)])
def test_head_ce_backward_many_tiles_and_chunks(M, Z):
    """More X tiles than SMs (a CTA owns several tiles: every mbarrier wraps) in either pass, long Y streams in the other."""
    import ananke_abm_b200 as ab
    dev = _cuda()
    g = torch.Generator().manual_seed(M * 7 + Z)
    emb = torch.randn(M, 64, generator=g).to(dev).requires_grad_(True)
    table = torch.randn(Z, 64, generator=g).to(dev).requires_grad_(True)
    tgt = torch.randint(0, Z, (M,), generator=g).to(dev)
    w = torch.rand(M, generator=g).to(dev)
    (ab.head_ce_rows(emb, table, tgt, 0.2) * w).sum().backward()
    e2, t2 = emb.detach().double().requires_grad_(True), table.detach().double().requires_grad_(True)
    (F.cross_entropy(_ref_logits(e2, t2, 0.2), tgt, reduction="none") * w.double()).sum().backward()
    # stated tolerance of the tensor-core path: logits come from 2-term bf16 splits (~2^-16 relative), and a gradient row
    # sums up to 20,000 such terms with cancellation (measured 1e-5 .. 1.2e-4 of the largest entry)
    for a, b in ((emb.grad, e2.grad), (table.grad, t2.grad)):
        assert float((a.double() - b).abs().max()) < 3e-4 * float(b.abs().max())
    # deterministic: a second run gives the same bits
    e3, t3 = emb.detach().clone().requires_grad_(True), table.detach().clone().requires_grad_(True)
    (ab.head_ce_rows(e3, t3, tgt, 0.2) * w).sum().backward()
    assert torch.equal(e3.grad, emb.grad) and torch.equal(t3.grad, table.grad)


def test_training_step_with_fused_ce_reaches_every_parameter():
    """End to end: GAT zone tables -> y0 -> dopri5 on the stage path -> decoder -> fused CE at the ground-truth snaps;
    one backward fills the gradients of the GAT layers, the context encoder, the drift net and the decoder, and the loss
    equals the reference formula evaluated on materialised logits (small Z)."""
    import ananke_abm_b200 as ab
    from ananke_abm_b200.graph import synthetic_zone_graph
    dev = _cuda()
    torch.manual_seed(3)
    mc = ab.ModeSepConfig()
    mc.precision, mc.ode_method = "bf16", "dopri5"
    Z, B, T = 96, 200, 6
    model = ab.GATODEModel(7, mc, heads=4).to(dev)
    ei, feats = synthetic_zone_graph(Z, k=6, seed=1)
    csr = ab.build_zone_csr(ei, Z).to(dev)
    g = torch.Generator().manual_seed(4)
    home, work = torch.randint(0, Z, (B,), generator=g).to(dev), torch.randint(0, Z, (B,), generator=g).to(dev)
    traits = torch.rand(B, 2, generator=g).to(dev)
    y_union = torch.randint(0, Z, (B, T), generator=g).to(dev)
    mask = (torch.rand(B, T, generator=g) < 0.4).to(dev)
    table, zemb = model.zone_tables(feats.to(dev), csr)
    y0 = model.initial_state(table, zemb, home, work, traits)
    y_path = model.integrate(y0, torch.linspace(0.0, 3.0, T, device=dev))
    pred_emb = model.decoder(y_path[:, :, :mc.emb_dim].permute(1, 0, 2))
    loss = ab.ce_at_snaps_fused(pred_emb, table, y_union, mask, mc.softmax_tau)
    loss.backward()
    logits = _ref_logits(pred_emb.detach().double(), table.detach().double(), mc.softmax_tau)
    ref = F.cross_entropy(logits[mask], y_union[mask], reduction="mean")
    assert abs(float(loss.detach()) - float(ref)) < 1e-5 * abs(float(ref))
    missing = [n for n, p in model.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all() or float(p.grad.abs().sum()) == 0.0]
    assert not missing, missing


@pytest.mark.parametrize("B,T,Z", [(37, 9, 500), (400, 12, 2000)])
def test_ce_and_expected_distance_fused_match_reference_formulas(B, T, Z):
    """ce_at_snaps + expected_distance_at_snaps (losses.py:14-22, 34-44) from one sweep: values and gradients of a
    weighted sum of both against float64 autograd through materialised logits."""
    import ananke_abm_b200 as ab
    dev = _cuda()
    g = torch.Generator().manual_seed(B + Z)
    emb = torch.randn(B, T, 64, generator=g).to(dev).requires_grad_(True)
    table = torch.randn(Z, 64, generator=g).to(dev).requires_grad_(True)
    xy = torch.rand(Z, 2, generator=g)
    dist = torch.cdist(xy, xy).to(dev)                      # [Z, Z] like the reference's dist_mat
    y = torch.randint(0, Z, (B, T), generator=g).to(dev)
    mask = (torch.rand(B, T, generator=g) < 0.35).to(dev)
    y = torch.where(mask, y, torch.full_like(y, -1))
    ce, ed = ab.ce_and_expected_distance_at_snaps_fused(emb, table, y, dist, mask, 0.2)
    (1.0 * ce + 0.5 * ed).backward()                        # w_ce, w_dist of mode_sep/config.py:39-41
    e2, t2 = emb.detach().double().requires_grad_(True), table.detach().double().requires_grad_(True)
    logits = _ref_logits(e2, t2, 0.2)
    ref_ce = F.cross_entropy(logits[mask], y[mask], reduction="mean")
    probs = torch.softmax(logits, dim=-1)
    ref_ed = (dist.double()[y.clamp(min=0)] * probs).sum(-1)[mask].mean()
    (1.0 * ref_ce + 0.5 * ref_ed).backward()
    assert abs(float(ce.detach()) - float(ref_ce)) < 1e-5 * abs(float(ref_ce))
    assert abs(float(ed.detach()) - float(ref_ed)) < 2e-5 * abs(float(ref_ed))
    for a, b in ((emb.grad, e2.grad), (table.grad, t2.grad)):
        assert float((a.double() - b).abs().max()) < 1e-4 * float(b.abs().max())


def test_fused_losses_match_the_unmodified_reference_on_the_fixture():
    """tests/golden/loss_terms_fixture.npz holds ce_at_snaps / expected_distance_at_snaps evaluated by the UNMODIFIED
    reference (mode_sep/architecture/losses.py) on the frozen fixture logits; the fused head must reproduce them from
    (pred_emb, class_table) alone -- at the GT snaps and on the stay-aux mask (mode_sep/train/train.py:131-133)."""
    import numpy as np
    from pathlib import Path
    import ananke_abm_b200 as ab
    dev = _cuda()
    gd = Path(__file__).parent / "golden"
    g, lt = np.load(gd / "mode_sep_fixture.npz"), np.load(gd / "loss_terms_fixture.npz")
    pred_emb = torch.from_numpy(g["pred_emb"]).to(dev)
    table = torch.from_numpy(g["sd_class_table"]).to(dev)
    dist = torch.from_numpy(g["dist_mat"]).float().to(dev)
    for y_key, m_key, ce_key, d_key in (("y_union", "ub_is_gt_union", "ce_gt", "dist_gt"),
                                        ("ub_stay_loc_ids", "ub_stay_non_gt_mask", "ce_aux", "dist_aux")):
        y, mask = torch.from_numpy(g[y_key]).to(dev), torch.from_numpy(g[m_key]).to(dev)
        ce, ed = ab.ce_and_expected_distance_at_snaps_fused(pred_emb, table, y, dist, mask, 0.2)
        assert abs(float(ce) - float(lt[ce_key])) < 1e-5 * abs(float(lt[ce_key])), (ce_key, float(ce), float(lt[ce_key]))
        assert abs(float(ed) - float(lt[d_key])) < 2e-5 * abs(float(lt[d_key])), (d_key, float(ed), float(lt[d_key]))
        assert abs(float(ab.ce_at_snaps_fused(pred_emb, table, y, mask, 0.2)) - float(lt[ce_key])) < 1e-5 * abs(float(lt[ce_key]))
