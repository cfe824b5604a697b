"""Fused classification losses of the mode_sep head.

`ce_at_snaps` of the reference (mode_sep/architecture/losses.py:14-22) takes the `[B, T, Z]` logits, which cannot exist at
configs[2] scale (3.9 TB).  `ce_at_snaps_fused` takes what the logits are made of -- `pred_emb [B, T, E]` and
`class_table [Z, E]` (mode_sep/architecture/model.py:196-199: cosine similarity / tau) -- and returns the same scalar:
the forward streams the zone table through the tensor cores (`ab200_head_ce_forward`: per-row log-sum-exp + target logit,
no logits in HBM); the backward (`ab200_head_ce_backward`) recomputes the probabilities tile by tile on the tensor
cores and consumes them on chip.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class _HeadLosses(torch.autograd.Function):
    """per-row cross entropy (and, with `dist_mat`, expected distance) of cos(emb, table) / tau against `target`."""

    @staticmethod
    def forward(ctx, emb, table, target, tau: float, dist_mat):
        L = _lib.lib()
        if not emb.is_cuda:
            raise _lib.Ab200Error("pred_emb must be a CUDA tensor: ananke_abm_b200 has no CPU path")
        M, E = emb.shape
        Z = table.shape[0]
        embc, tablec = emb.detach().contiguous().float(), table.detach().contiguous().float()
        tgt = target.contiguous().to(torch.int64)
        nbytes = L.ab200_head_workspace_bytes(Z, E)
        if nbytes == 0:
            raise _lib.Ab200Error("fused head is instantiated for emb_dim = 64")
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=emb.device)
        lse = torch.empty(M, dtype=torch.float32, device=emb.device)
        tl = torch.empty(M, dtype=torch.float32, device=emb.device)
        dm = None if dist_mat is None else dist_mat.detach().contiguous().float()
        if dm is not None and dm.shape != (Z, Z):
            raise ValueError("dist_mat must be [Z, Z]")
        ed = None if dm is None else torch.empty(M, dtype=torch.float32, device=emb.device)
        rc = L.ab200_head_ce_forward(embc.data_ptr(), tablec.data_ptr(), tgt.data_ptr(), M, Z, E, float(tau), lse.data_ptr(),
                                     tl.data_ptr(), None, None if dm is None else dm.data_ptr(), None if ed is None else ed.data_ptr(),
                                     ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "ab200_head_ce_forward")
        ctx.save_for_backward(embc, tablec, tgt, lse, *(() if dm is None else (dm, ed)))
        ctx.tau, ctx.has_dist = float(tau), dm is not None
        ce = lse - tl
        if dm is None:
            ctx.mark_non_differentiable()
            return ce, torch.zeros_like(ce)
        return ce, ed

    @staticmethod
    def backward(ctx, g_rows, g_dist):
        # d loss_m / d logit_mz = g_m (softmax_mz - [z = y_m]) + g2_m softmax_mz (D[y_m, z] - E_m): recomputed tile by tile on
        # the tensor cores from the saved log-sum-exp (`ab200_head_ce_backward`: rows-outer pass -> d emb^, zones-outer
        # pass -> d table^, no atomics)
        L = _lib.lib()
        saved = ctx.saved_tensors
        emb, table, tgt, lse = saved[:4]
        dm, ed = (saved[4], saved[5]) if ctx.has_dist else (None, None)
        M, E = emb.shape
        Z = table.shape[0]
        g = g_rows.contiguous().float()
        g2 = g_dist.contiguous().float() if ctx.has_dist else None
        g_eh = torch.empty_like(emb)
        g_th = torch.empty_like(table)
        ws = torch.empty(int(L.ab200_head_ce_backward_workspace_bytes(M, Z, E)), dtype=torch.uint8, device=emb.device)
        stream = torch.cuda.current_stream().cuda_stream
        rc = L.ab200_head_ce_backward(emb.data_ptr(), table.data_ptr(), tgt.data_ptr(), lse.data_ptr(), g.data_ptr(),
                                      None if g2 is None else g2.data_ptr(), None if ed is None else ed.data_ptr(),
                                      None if dm is None else dm.data_ptr(), M, Z, E, ctx.tau,
                                      g_eh.data_ptr(), g_th.data_ptr(), ws.data_ptr(), ws.numel(), stream)
        _lib.check(rc, "ab200_head_ce_backward")
        st = C.c_int32(0)
        _lib.check(L.ab200_head_ce_backward_status(ws.data_ptr(), M, Z, C.byref(st), stream), "ab200_head_ce_backward_status")
        if st.value:
            raise _lib.Ab200Error(f"head cross-entropy backward: barrier timeout (status {st.value})")
        # through y = x / (n + eps), n = |x|:  g_x = g_y / (n + eps) - y (g_y . y) / n
        n_e, n_t = emb.norm(dim=-1, keepdim=True), table.norm(dim=-1, keepdim=True)
        eh, th = emb / (n_e + 1e-8), table / (n_t + 1e-8)
        g_emb = g_eh / (n_e + 1e-8) - eh * ((g_eh * eh).sum(-1, keepdim=True) / n_e.clamp_min(1e-30))
        g_table = g_th / (n_t + 1e-8) - th * ((g_th * th).sum(-1, keepdim=True) / n_t.clamp_min(1e-30))
        return g_emb, g_table, None, None, None


def head_loss_rows(pred_emb: torch.Tensor, class_table: torch.Tensor, target: torch.Tensor, tau: float = 0.2,
                   dist_mat: torch.Tensor = None):
    """-> (cross entropy, expected distance) per row of `pred_emb [..., E]` against `target [...]` in ONE sweep over the
    zones; expected distance = sum_z softmax_z * dist_mat[target, z] (zeros when `dist_mat` is None).  Rows whose target is
    outside [0, Z) are scored against zone 0 (mask them).  With `dist_mat` the rows are processed in target order so that a
    tile of 128 rows reads one or two lines of the [Z, Z] matrix instead of 128."""
    lead = pred_emb.shape[:-1]
    emb = pred_emb.reshape(-1, pred_emb.shape[-1])
    tgt = target.reshape(-1)
    if dist_mat is None:
        ce, ed = _HeadLosses.apply(emb, class_table, tgt, tau, None)
        return ce.view(lead), ed.view(lead)
    order = torch.argsort(tgt, stable=True)
    inv = torch.empty_like(order)
    inv[order] = torch.arange(order.numel(), device=order.device)
    ce, ed = _HeadLosses.apply(emb[order], class_table, tgt[order], tau, dist_mat)
    return ce[inv].view(lead), ed[inv].view(lead)


def head_ce_rows(pred_emb: torch.Tensor, class_table: torch.Tensor, target: torch.Tensor, tau: float = 0.2) -> torch.Tensor:
    """Per-row cross entropy `-log softmax(cos(pred_emb, class_table) / tau)[target]` for `pred_emb [..., E]`,
    `target [...]`; rows whose target is outside [0, Z) are scored against zone 0 (mask them)."""
    return head_loss_rows(pred_emb, class_table, target, tau, None)[0]


def ce_at_snaps_fused(pred_emb: torch.Tensor, class_table: torch.Tensor, y_union: torch.Tensor, is_gt_mask: torch.Tensor,
                      tau: float = 0.2) -> torch.Tensor:
    """`ce_at_snaps(model.head(...)[1], y_union, is_gt_mask)` of the reference (losses.py:14-22) from the head's inputs:
    mean cross entropy over the (agent, time) pairs where `is_gt_mask` is set; 0 when the mask is empty."""
    mask = is_gt_mask
    if int(mask.sum()) == 0:
        return torch.tensor(0.0, dtype=pred_emb.dtype, device=pred_emb.device)
    rows = head_ce_rows(pred_emb[mask], class_table, y_union[mask], tau)
    return rows.mean()


def ce_and_expected_distance_at_snaps_fused(pred_emb: torch.Tensor, class_table: torch.Tensor, y_union: torch.Tensor,
                                            dist_mat: torch.Tensor, is_gt_mask: torch.Tensor, tau: float = 0.2):
    """(`ce_at_snaps`, `expected_distance_at_snaps`) of the reference (losses.py:14-22, 34-44) from the head's inputs in
    one pass over the zones: mean over the masked (agent, time) pairs of the cross entropy and of
    sum_z softmax(logits)_z * dist_mat[y, z]; (0, 0) when the mask is empty."""
    mask = is_gt_mask
    if int(mask.sum()) == 0:
        z = torch.tensor(0.0, dtype=pred_emb.dtype, device=pred_emb.device)
        return z, z.clone()
    ce, ed = head_loss_rows(pred_emb[mask], class_table, y_union[mask], tau, dist_mat)
    return ce.mean(), ed.mean()


# --------------------------------------------------------------------------------------------------------------------
# embedding-space terms (mse, travel margin / monotonicity, velocity regularisers) and the whole mode_sep objective
# --------------------------------------------------------------------------------------------------------------------
def _strided_rows(x: torch.Tensor):
    """[B, T, E] fp32 with unit inner stride and 16-byte aligned rows -> (tensor, stride_b, stride_t); copies otherwise"""
    if x.dtype != torch.float32 or x.stride(-1) != 1 or x.stride(0) % 4 or x.stride(1) % 4 or x.data_ptr() % 16:
        x = x.float().contiguous()
    return x, int(x.stride(0)), int(x.stride(1))


class _EmbLosses(torch.autograd.Function):
    """-> the six embedding-space loss terms of the mode_sep objective as 0-dim tensors, from ONE pass over (pred_emb, v_t):
    [mse at snaps, mse inside stays, travel margin, travel monotonicity, stay velocity, move velocity]
    (`ab200_emb_losses_forward` / `_backward`; masked means with the reference's "0 when the mask is empty" rule)."""

    @staticmethod
    def forward(ctx, pred_emb, v_t, class_table, y_union, is_gt, y_stay, stay_non_gt, travel_mask, prev_idx, dest_idx, gt_interior,
                m_travel: float, epsilon_mono: float, v_min: float, v_max: float):
        L = _lib.lib()
        if not pred_emb.is_cuda:
            raise _lib.Ab200Error("pred_emb must be a CUDA tensor: ananke_abm_b200 has no CPU path")
        B, T, E = pred_emb.shape
        Z = class_table.shape[0]
        emb, esb, est = _strided_rows(pred_emb.detach())
        vt, vsb, vst = _strided_rows(v_t.detach())
        table = class_table.detach().contiguous().float()
        idx = [t.contiguous().to(torch.int64) for t in (y_union, y_stay, prev_idx, dest_idx)]
        msk = [t.contiguous().to(torch.bool) for t in (is_gt, stay_non_gt, travel_mask, gt_interior)]
        sums = torch.zeros(12, dtype=torch.float64, device=pred_emb.device)
        par = (float(m_travel), float(epsilon_mono), float(v_min), float(v_max))
        rc = L.ab200_emb_losses_forward(emb.data_ptr(), esb, est, vt.data_ptr(), vsb, vst, table.data_ptr(), idx[0].data_ptr(),
                                        msk[0].data_ptr(), idx[1].data_ptr(), msk[1].data_ptr(), msk[2].data_ptr(), idx[2].data_ptr(),
                                        idx[3].data_ptr(), msk[3].data_ptr(), B, T, E, Z, *par, sums.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "ab200_emb_losses_forward")
        cnt = torch.stack([sums[1], sums[3], sums[5], sums[8], sums[3], sums[11]])
        num = torch.stack([sums[0], sums[2], sums[4], 0.5 * (sums[6] + sums[7]), sums[9], sums[10]])
        terms = torch.where(cnt > 0, num / cnt.clamp_min(1.0), torch.zeros_like(num)).float()
        ctx.save_for_backward(emb, vt, table, *idx, *msk, cnt)
        ctx.par, ctx.strides, ctx.shape = par, (esb, est, vsb, vst), (B, T, E, Z)
        return tuple(terms.unbind(0))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *g):
        L = _lib.lib()
        emb, vt, table, y_union, y_stay, prev_idx, dest_idx, is_gt, stay_non_gt, travel_mask, gt_interior, cnt = ctx.saved_tensors
        B, T, E, Z = ctx.shape
        esb, est, vsb, vst = ctx.strides
        gs = torch.stack([x.reshape(()).double() for x in g])
        scale = torch.tensor([1.0, 1.0, 1.0, 0.5, 1.0, 1.0], dtype=torch.float64, device=emb.device)
        coef = torch.where(cnt > 0, gs * scale / cnt.clamp_min(1.0), torch.zeros_like(gs)).float().contiguous()
        d_emb = torch.empty((B, T, E), dtype=torch.float32, device=emb.device)
        d_v = torch.empty((B, T, E), dtype=torch.float32, device=emb.device)
        d_table = torch.zeros((Z, E), dtype=torch.float32, device=emb.device)
        rc = L.ab200_emb_losses_backward(emb.data_ptr(), esb, est, vt.data_ptr(), vsb, vst, table.data_ptr(), y_union.data_ptr(),
                                         is_gt.data_ptr(), y_stay.data_ptr(), stay_non_gt.data_ptr(), travel_mask.data_ptr(),
                                         prev_idx.data_ptr(), dest_idx.data_ptr(), gt_interior.data_ptr(), B, T, E, Z, *ctx.par,
                                         coef.data_ptr(), d_emb.data_ptr(), d_v.data_ptr(), d_table.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "ab200_emb_losses_backward")
        return (d_emb, d_v, d_table) + (None,) * 12


def emb_loss_terms(pred_emb, v_t, class_table, y_union, is_gt, y_stay, stay_non_gt, travel_mask, prev_idx, dest_idx, gt_interior,
                   m_travel: float = 0.10, epsilon_mono: float = 0.01, v_min_move: float = 0.2, v_max_move: float = 1.0):
    """The reference's `mse_at_snaps` (at snaps and, second value, at the non-snap stay points), `travel_margin_loss`,
    `travel_monotonicity_loss` (mode_sep/architecture/losses.py:24-31, 56-115) and the two velocity regularisers of the
    training loop (mode_sep/train/train.py:137-153) in one fused pass; same values, same "0 when the mask is empty" rule."""
    names = ("mse", "stay_mse", "travel_margin", "travel_mono", "stay_vel", "move_vel")
    vals = _EmbLosses.apply(pred_emb, v_t, class_table, y_union, is_gt, y_stay, stay_non_gt, travel_mask, prev_idx, dest_idx,
                            gt_interior, m_travel, epsilon_mono, v_min_move, v_max_move)
    return dict(zip(names, vals))


def mode_sep_total_loss(config, pred_emb, v_t, class_table, union, y_union, dist_mat):
    """The COMPLETE training objective of the reference's mode_sep loop (mode_sep/train/train.py:111-159), from the head's
    inputs instead of the `[B, T, Z]` logits:

        total_loss(...)                              losses.py:118-158   w_ce ce + w_mse mse + w_dist dist + travel terms
      + w_stay_aux (ce + mse + dist at stay_non_gt)  train.py:121-135
      + w_stay_vel_core stay_vel + w_move_vel_hinge move_vel          train.py:137-159

    `union` is the `UnionBatch` of `build_union_batch`, `y_union [B, T]` the zone id at snaps (-1 elsewhere, train.py:102-108).
    Cross entropy and expected distance come from the fused tensor-core head (two masked row sets), every embedding-space
    term from the fused pass above.  -> (total, parts) with `parts` a dict of 0-dim tensors (no host sync)."""
    tau = config.softmax_tau
    ce, dist = ce_and_expected_distance_at_snaps_fused(pred_emb, class_table, y_union, dist_mat, union.is_gt_union, tau)
    aux_ce, aux_dist = ce_and_expected_distance_at_snaps_fused(pred_emb, class_table, union.stay_loc_ids, dist_mat,
                                                               union.stay_non_gt_mask, tau)
    e = emb_loss_terms(pred_emb, v_t, class_table, y_union, union.is_gt_union, union.stay_loc_ids, union.stay_non_gt_mask,
                       union.travel_mask, union.prev_zone_idx, union.dest_zone_idx, union.gt_interior_mask,
                       config.m_travel, config.epsilon_mono, config.v_min_move, config.v_max_move)
    base = (config.w_ce * ce + config.w_mse * e["mse"] + config.w_dist * dist + config.w_travel_margin * e["travel_margin"]
            + config.w_travel_mono * e["travel_mono"])
    aux = config.w_stay_aux * (aux_ce + e["stay_mse"] + aux_dist)
    total = base + aux + config.w_stay_vel_core * e["stay_vel"] + config.w_move_vel_hinge * e["move_vel"]
    parts = {"ce": ce, "mse": e["mse"], "dist": dist, "travel_margin": e["travel_margin"], "travel_mono": e["travel_mono"],
             "stay_aux": aux, "stay_vel": e["stay_vel"], "move_vel": e["move_vel"], "base": base}
    return total, parts
