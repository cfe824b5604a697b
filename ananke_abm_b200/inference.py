"""`inference` -- predicted activity/zone labels for batches of agents (README namespace, /root/reference/README.md:57).

Mirrors what the reference's inference paths compute from a forward pass: `logits.argmax(-1)`
(/root/reference/src/ananke_abm/models/mode_sep/inference/inference.py:57,63;
 /root/reference/src/ananke_abm/models/latent_ode/inference/inference.py:200-202), but batched over agents and
chunked over time so that the `[B, T, Z]` logits tensor is never materialised at once (at 1M agents x 97 x 10k zones
it would be 3.9 TB -- SURVEY.md §7.3).
"""
from __future__ import annotations

from typing import Optional

import torch


@torch.no_grad()
def predict_labels(model, times_union, home_idx, work_idx, person_traits_raw, zone_features=None, graph=None,
                   t_chunk: int = 8, agent_chunk: int = 262_144) -> torch.Tensor:
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
        for ts in range(0, T, t_chunk):
            seg = y_path[ts:ts + t_chunk]
            logits = model.head(seg, class_table)[1] if gat else model.head(seg)[1]
            out[sl, ts:ts + t_chunk] = logits.argmax(-1)
    return out
