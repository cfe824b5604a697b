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


# ---- embedding-space terms (mse, travel margin / monotonicity, velocity regularisers) and the whole objective ---------------------
def _ref_emb_terms(pred, v, table, y_union, is_gt, y_stay, m_stay, travel, prev, dest, m_move, m_travel=0.10, eps=0.01, vmin=0.2, vmax=1.0):
    """the reference's own expressions (mode_sep/architecture/losses.py:24-31, 46-115; mode_sep/train/train.py:137-153)"""
    z = pred.new_zeros(())

    def mse(y, mask):
        return (pred - table[y.clamp(min=0)]).pow(2).sum(-1)[mask].mean() if mask.any() else z

    def d2c(idx):
        return (pred - table[idx.clamp_min(0)]).pow(2).sum(-1).sqrt()
    out = {"mse": mse(y_union, is_gt), "stay_mse": mse(y_stay, m_stay), "travel_margin": z, "travel_mono": z}
    if travel.any():
        dp, dd = d2c(prev), d2c(dest)
        out["travel_margin"] = (m_travel - (dp - dd))[travel].clamp(min=0.0).mean()
        pair = travel[:, :-1] & travel[:, 1:] & (prev[:, :-1] == prev[:, 1:]) & (dest[:, :-1] == dest[:, 1:])
        if pair.any():
            away = (dp[:, :-1][pair] - dp[:, 1:][pair] + eps).clamp(min=0.0)
            toward = (dd[:, 1:][pair] - dd[:, :-1][pair] + eps).clamp(min=0.0)
            out["travel_mono"] = (away.mean() + toward.mean()) * 0.5
    v_abs = v.norm(dim=-1)
    out["stay_vel"] = (v_abs[m_stay] ** 2).mean() if m_stay.any() else z
    if m_move.any():
        vm = v_abs[m_move]
        out["move_vel"] = ((vmin - vm).clamp(min=0.0) ** 2 + (vm - vmax).clamp(min=0.0) ** 2).mean()
    else:
        out["move_vel"] = z
    return out


def _random_union(B, T, Z, g, dev):
    """random stay / travel structure per agent with the UnionBatch field semantics (batching.py:15-28)"""
    is_gt = torch.zeros(B, T, dtype=torch.bool)
    stay = torch.zeros(B, T, dtype=torch.bool)
    travel = torch.zeros(B, T, dtype=torch.bool)
    y_union = torch.full((B, T), -1, dtype=torch.long)
    y_stay = torch.full((B, T), -1, dtype=torch.long)
    prev = torch.full((B, T), -1, dtype=torch.long)
    dest = torch.full((B, T), -1, dtype=torch.long)
    for b in range(B):
        t, zone = 0, int(torch.randint(0, Z, (1,), generator=g))
        while t < T:
            n = int(torch.randint(1, 6, (1,), generator=g))
            if int(torch.randint(0, 2, (1,), generator=g)):          # a stay: snaps at both ends
                seg = slice(t, min(T, t + n))
                stay[b, seg] = True
                y_stay[b, seg] = zone
                is_gt[b, t] = True
                y_union[b, t] = zone
            else:                                                      # a travel leg to a new zone
                nz = int(torch.randint(0, Z, (1,), generator=g))
                seg = slice(t, min(T, t + n))
                travel[b, seg] = True
                prev[b, seg], dest[b, seg] = zone, nz
                zone = nz
            t += n
    stay_non_gt = stay & ~is_gt
    gt_idx = is_gt.clone()
    first = is_gt.float().cumsum(1) == 1
    last = is_gt.flip(1).float().cumsum(1).flip(1) == 1
    gt_interior = is_gt & ~(first & is_gt) & ~(last & is_gt)
    del gt_idx
    return [x.to(dev) for x in (y_union, is_gt, y_stay, stay_non_gt, travel, prev, dest, gt_interior)]


@pytest.mark.parametrize("B,T,Z", [(1, 1, 8), (3, 7, 8), (130, 29, 500), (257, 97, 10_000)])
def test_emb_loss_terms_match_the_reference_expressions(B, T, Z):
    import ananke_abm_b200 as ab
    dev = _cuda()
    g = torch.Generator().manual_seed(B * 1000 + T)
    table = (0.3 * torch.randn(Z, 64, generator=g)).to(dev).requires_grad_(True)
    y_path = (0.3 * torch.randn(T, B, 160, generator=g)).to(dev).requires_grad_(True)      # the solver's layout: v_t is a strided view
    yb = y_path.permute(1, 0, 2)
    v_t = yb[:, :, 64:128]
    pred = (yb[:, :, :64] * 1.0 + 0.1).contiguous()
    pred.retain_grad()
    fields = _random_union(B, T, Z, g, dev)
    terms = ab.emb_loss_terms(pred, v_t, table, *fields)
    w = {"mse": 0.7, "stay_mse": 1.3, "travel_margin": 2.1, "travel_mono": 0.9, "stay_vel": 5.0, "move_vel": 1.1}
    sum(w[k] * terms[k] for k in w).backward()
    p2, v2, t2 = pred.detach().double().requires_grad_(True), v_t.detach().double().requires_grad_(True), table.detach().double().requires_grad_(True)
    ref = _ref_emb_terms(p2, v2, t2, *fields)
    sum(w[k] * ref[k] for k in w).backward()
    for k in w:
        assert abs(float(terms[k]) - float(ref[k])) <= 2e-6 * max(abs(float(ref[k])), 1e-3), (k, float(terms[k]), float(ref[k]))

    def rel(a, b):
        b = torch.zeros_like(a, dtype=torch.float64) if b is None else b        # a term whose mask is empty leaves no gradient
        return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-12))
    assert rel(pred.grad, p2.grad) < 1e-5
    assert rel(y_path.grad.permute(1, 0, 2)[:, :, 64:128], v2.grad) < 1e-5
    assert rel(table.grad, t2.grad) < 2e-5


def test_mode_sep_full_objective_fused_matches_the_unmodified_reference(golden_mode_sep):
    """The COMPLETE training objective of mode_sep/train/train.py:111-159 through `mode_sep_total_loss` (fused head for cross
    entropy / expected distance, fused pass for every embedding-space term, no [B, T, Z] logits): the loss and EVERY parameter
    gradient equal the vectors minted from the unmodified reference's training step (tests/golden/make_golden.py)."""
    from types import SimpleNamespace
    import ananke_abm_b200 as ab
    dev = _cuda()
    g = golden_mode_sep
    cfg = ab.ModeSepConfig()
    m = ab.ModeSepModel(8, cfg)
    m.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd_")}, strict=True)
    m = m.to(dev)
    tt = lambda k: torch.from_numpy(g[k]).to(dev)   # noqa: E731
    union = SimpleNamespace(**{f: tt("ub_" + f) for f in ("is_gt_union", "stay_loc_ids", "stay_non_gt_mask", "travel_mask",
                                                          "prev_zone_idx", "dest_zone_idx", "gt_interior_mask")})
    y0 = m.initial_state(tt("home_idx"), tt("work_idx"), tt("traits"))
    y_path = m.integrate(y0, tt("times_union"))
    yb = y_path.permute(1, 0, 2)
    pred_emb = m.decoder(yb[:, :, :64])
    total, parts = ab.mode_sep_total_loss(cfg, pred_emb, yb[:, :, 64:128], m.class_table, union, tt("y_union"), tt("dist_mat"))
    assert abs(float(total) - float(g["loss_total"])) < 2e-5 * abs(float(g["loss_total"]))
    assert abs(float(parts["base"]) - float(g["loss_base"])) < 2e-5 * abs(float(g["loss_base"]))
    total.backward()
    for name, p in m.named_parameters():
        ref = torch.from_numpy(g["grad_" + name])
        got = p.grad.cpu() if p.grad is not None else torch.zeros_like(ref)
        scale = ref.abs().max().clamp_min(1e-12)
        assert float((got - ref).abs().max() / scale) < 1e-4, (name, float((got - ref).abs().max() / scale))
