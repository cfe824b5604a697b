"""`inference` -- predicted activity/zone labels for batches of agents (README namespace, /root/reference/README.md:57).

Mirrors what the reference's inference paths compute from a forward pass: `logits.argmax(-1)`
(/root/reference/src/ananke_abm/models/mode_sep/inference/inference.py:57,63;
 /root/reference/src/ananke_abm/models/latent_ode/inference/inference.py:200-202), but batched over agents and
chunked over time so that the `[B, T, Z]` logits tensor is never materialised at once (at 1M agents x 97 x 10k zones
it would be 3.9 TB -- SURVEY.md §7.3).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib


def head_argmax(pred_emb: torch.Tensor, class_table: torch.Tensor, tau: float = 0.2, return_logit: bool = False):
    """labels = argmax_z cos(pred_emb, class_table[z]) / tau for every row of `pred_emb[..., E]`, fused on the tensor cores
    (`ab200_head_argmax`): the [.., Z] logits are never materialised.  Equals
    `model.head(...)[1].argmax(-1)` of the reference (mode_sep/architecture/model.py:196-199 + inference.py:57)."""
    L = _lib.lib()
    if not pred_emb.is_cuda:
        raise _lib.Ab200Error("pred_emb must be a CUDA tensor: ananke_abm_b200 has no CPU path")
    E = pred_emb.shape[-1]
    lead = pred_emb.shape[:-1]
    emb = pred_emb.detach().reshape(-1, E).contiguous().float()
    table = class_table.detach().contiguous().float()
    M, Z = emb.shape[0], table.shape[0]
    nbytes = L.ab200_head_workspace_bytes(Z, E)
    if nbytes == 0:
        raise _lib.Ab200Error("fused head is instantiated for emb_dim = 64")
    ws = torch.empty(int(nbytes), dtype=torch.uint8, device=emb.device)
    labels = torch.empty(M, dtype=torch.int64, device=emb.device)
    best = torch.empty(M, dtype=torch.float32, device=emb.device) if return_logit else None
    rc = L.ab200_head_argmax(emb.data_ptr(), table.data_ptr(), M, Z, E, float(tau), labels.data_ptr(),
                             None if best is None else best.data_ptr(), ws.data_ptr(), ws.numel(),
                             torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ab200_head_argmax")
    labels = labels.view(lead)
    return (labels, best.view(lead)) if return_logit else labels


@torch.no_grad()
def predict_labels(model, times_union, home_idx, work_idx, person_traits_raw, zone_features=None, graph=None,
                   t_chunk: int = 8, agent_chunk: int = 262_144, fused: bool = True) -> torch.Tensor:
    """-> int64 labels [B, T]: argmax zone per agent and time point; identical indexing to the reference
    (`labels[b, t]` belongs to agent `b` of the input batch at `times_union[t]`)."""
    B, T = home_idx.shape[0], times_union.shape[0]
    gat = zone_features is not None
    if gat:
        class_table, zone_embed = model.zone_tables(zone_features, graph)
    out = torch.empty((B, T), dtype=torch.int64, device=home_idx.device)
    for s in range(0, B, agent_chunk):
        sl = slice(s, min(B, s + agent_chunk))
        if gat:
            y0 = model.initial_state(class_table, zone_embed, home_idx[sl], work_idx[sl], person_traits_raw[sl])
        else:
            y0 = model.initial_state(home_idx[sl], work_idx[sl], person_traits_raw[sl])
        y_path = model.integrate(y0, times_union)
        E = model.config.emb_dim
        table = class_table if gat else model.class_table
        if fused and E == 64:
            for ts in range(0, T, t_chunk):
                pred_emb = model.decoder(y_path[ts:ts + t_chunk, :, :E].permute(1, 0, 2))      # [b, t, E]: decoder MLP (library GEMMs)
                out[sl, ts:ts + t_chunk] = head_argmax(pred_emb, table, model.config.softmax_tau)
        else:
            for ts in range(0, T, t_chunk):
                seg = y_path[ts:ts + t_chunk]
                logits = model.head(seg, class_table)[1] if gat else model.head(seg)[1]
                out[sl, ts:ts + t_chunk] = logits.argmax(-1)
    return out
