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
