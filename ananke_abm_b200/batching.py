"""Union-time-grid batching with the reference's batch layouts, tensorised over agents.

Two collate routines feed the ODE hot path in the reference; both are per-person / per-segment Python loops with
`.item()` calls, fine for B = 2 and unusable at 10^4..10^6 agents:

  build_union_batch            /root/reference/src/ananke_abm/models/mode_sep/data_process/batching.py:76-182
                               -> `UnionBatch` (same field names, dtypes, shapes and index conventions, :15-28)
  unify_and_interpolate_batch  /root/reference/src/ananke_abm/models/latent_ode/data_process/batching.py:12-128
                               -> the same 16-key collate dict

Here every [B, T] field is produced by broadcast / cumulative-scan / gather operations on whatever device the inputs
live on (agents are processed in chunks to bound the [B, T, S] intermediates); only the O(T) construction of the
shared time grid keeps the reference's scalar `torch.linspace` calls, so that `times_union` is bit-identical.
Agent `b` of the input list is row `b` of every output: batch indexing is unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch


@dataclass
class UnionBatch:                      # mode_sep/data_process/batching.py:15-28
    times_union: torch.Tensor          # (T,) float32
    is_gt_union: torch.Tensor          # (B, T) bool
    snap_indices: torch.Tensor         # (B, T) long, index into the person's loc_ids or -1
    stay_mask: torch.Tensor            # (B, T) bool
    gt_interior_mask: torch.Tensor     # (B, T) bool, GT snaps excluding first / last
    stay_non_gt_mask: torch.Tensor     # (B, T) bool
    stay_loc_ids: torch.Tensor         # (B, T) long or -1
    travel_mask: torch.Tensor          # (B, T) bool
    prev_zone_idx: torch.Tensor        # (B, T) long or -1
    dest_zone_idx: torch.Tensor        # (B, T) long or -1
    progress_s: torch.Tensor           # (B, T) float32 in [0, 1]
    min_dt: float


def insert_internal_points(sorted_times: torch.Tensor, K: int) -> torch.Tensor:
    """batching.py:31-47 -- K equally spaced interior points per gap (the reference's own scalar linspace calls:
    the grid has O(T) points and must match bit for bit)."""
    if sorted_times.numel() <= 1:
        return sorted_times
    pieces: List[torch.Tensor] = []
    for i in range(sorted_times.numel() - 1):
        t0, t1 = sorted_times[i], sorted_times[i + 1]
        pieces.append(t0.unsqueeze(0))
        if K > 0:
            internal = torch.linspace(float(t0), float(t1), steps=K + 2, dtype=sorted_times.dtype, device=sorted_times.device)[1:-1]
            if internal.numel() > 0:
                pieces.append(internal)
    pieces.append(sorted_times[-1:].clone())
    return torch.unique(torch.cat(pieces), sorted=True)


def _pad(rows: Sequence[torch.Tensor], fill, dtype, device) -> torch.Tensor:
    n = max((int(r.numel()) for r in rows), default=0)
    out = torch.full((len(rows), max(n, 1)), fill, dtype=dtype, device=device)
    for i, r in enumerate(rows):
        if r.numel():
            out[i, :r.numel()] = r.to(device=device, dtype=dtype)
    return out


def build_union_batch(persons: Sequence, config, device, chunk: int = 4096) -> UnionBatch:
    """`persons[i]` needs `.times_snap (S,) float32`, `.loc_ids (S,) long`, `.stay_segments [(t0, t1, loc)]` and optionally
    `.stay_intervals [(t0, t1)]` (defaults to the segments' intervals) -- the reference's `PersonData`
    (mode_sep/data_process/data.py:18-27); `config` needs `K_internal` and `time_match_tol`."""
    device = torch.device(device)
    all_times = [p.times_snap for p in persons if p.times_snap.numel() > 0]
    if not all_times:
        raise ValueError("No snap times found for any person in the batch.")
    times_union = torch.unique(torch.cat([t.cpu() for t in all_times]), sorted=True)
    times_union = insert_internal_points(times_union, config.K_internal).to(device)
    B, T = len(persons), int(times_union.shape[0])
    tol = float(config.time_match_tol)
    ar = torch.arange(T, device=device)

    outs = {k: [] for k in ("is_gt", "sidx", "stay", "interior", "stay_loc", "travel", "prev", "dest", "prog")}
    for c0 in range(0, B, chunk):
        ps = persons[c0:c0 + chunk]
        nb = len(ps)
        ts = _pad([p.times_snap for p in ps], float("inf"), times_union.dtype, device)                     # [b, S]
        loc = _pad([p.loc_ids for p in ps], -1, torch.long, device)                                          # [b, S]
        segs = [torch.as_tensor([[s[0], s[1], s[2]] for s in p.stay_segments], dtype=torch.float64).reshape(-1, 3) for p in ps]
        ivals = [torch.as_tensor([[s[0], s[1]] for s in getattr(p, "stay_intervals", None) or p.stay_segments],
                                 dtype=torch.float64).reshape(-1, 2) for p in ps]
        G = max(max((int(s.shape[0]) for s in segs), default=0), 1)
        GI = max(max((int(s.shape[0]) for s in ivals), default=0), 1)
        seg_t = torch.full((nb, G, 2), float("nan"), dtype=times_union.dtype, device=device)
        seg_loc = torch.full((nb, G), -1, dtype=torch.long, device=device)
        iv_t = torch.full((nb, GI, 2), float("nan"), dtype=times_union.dtype, device=device)
        for i, (s, v) in enumerate(zip(segs, ivals)):
            if s.shape[0]:
                seg_t[i, :s.shape[0]] = s[:, :2].to(device=device, dtype=times_union.dtype)    # float(t) -> grid dtype, as :69-70
                seg_loc[i, :s.shape[0]] = s[:, 2].to(device=device, dtype=torch.long)
            if v.shape[0]:
                iv_t[i, :v.shape[0]] = v.to(device=device, dtype=times_union.dtype)

        # (a) GT alignment: first snap within tol of each grid time (:50-63)
        eq = torch.isclose(times_union[None, :, None], ts[:, None, :], atol=tol, rtol=0.0)                  # [b, T, S]
        is_gt = eq.any(dim=2)
        sidx = torch.where(is_gt, eq.float().argmax(dim=2), torch.full((nb, T), -1, dtype=torch.long, device=device))
        # (b) stay mask (any interval, closed) and stay location (the LAST segment containing t wins, as the loop at :126-131)
        tu = times_union[None, :, None]
        stay = ((tu >= iv_t[:, None, :, 0]) & (tu <= iv_t[:, None, :, 1])).any(dim=2)
        in_seg = (tu >= seg_t[:, None, :, 0]) & (tu <= seg_t[:, None, :, 1])                                 # [b, T, G]
        last = (in_seg.long() * (torch.arange(G, device=device) + 1)).amax(dim=2)                           # 0 = none
        stay_loc = torch.where(last > 0, seg_loc.gather(1, (last - 1).clamp_min(0)), torch.full_like(last, -1))
        # (c) interior GT snaps: rank among the person's GT grid points, excluding first and last (:133-137)
        rank = is_gt.long().cumsum(dim=1)
        count = rank[:, -1:]
        interior = is_gt & (rank > 1) & (rank < count)
        # (e) travel metadata between consecutive GT snaps with different zones (:142-169)
        gt_pos = torch.where(is_gt, ar[None, :], torch.full((1, 1), -1, dtype=torch.long, device=device))
        j0 = gt_pos.cummax(dim=1).values                                                                    # last GT index <= t
        nxt = torch.where(is_gt, ar[None, :], torch.full((1, 1), T, dtype=torch.long, device=device))
        j1 = nxt.flip(1).cummin(dim=1).values.flip(1)                                                       # first GT index >= t
        inside = (~is_gt) & (j0 >= 0) & (j1 < T)
        j0c, j1c = j0.clamp_min(0), j1.clamp_max(T - 1)
        z0 = loc.gather(1, sidx.gather(1, j0c).clamp_min(0))
        z1 = loc.gather(1, sidx.gather(1, j1c).clamp_min(0))
        travel = inside & (z0 != z1)
        t0g, t1g = times_union[j0c], times_union[j1c]
        prog = ((times_union[None, :] - t0g) / (t1g - t0g).clamp(min=1e-8)).clamp(0.0, 1.0)
        neg = torch.full_like(z0, -1)
        outs["is_gt"].append(is_gt); outs["sidx"].append(sidx); outs["stay"].append(stay); outs["interior"].append(interior)
        outs["stay_loc"].append(stay_loc); outs["travel"].append(travel)
        outs["prev"].append(torch.where(travel, z0, neg)); outs["dest"].append(torch.where(travel, z1, neg))
        outs["prog"].append(torch.where(travel, prog, torch.zeros_like(prog)).to(torch.float32))

    cat = {k: torch.cat(v, dim=0) for k, v in outs.items()}
    diffs = times_union[1:] - times_union[:-1]
    min_dt = float(diffs.min().item()) if diffs.numel() > 0 else 1.0
    return UnionBatch(times_union=times_union, is_gt_union=cat["is_gt"], snap_indices=cat["sidx"], stay_mask=cat["stay"],
                      gt_interior_mask=cat["interior"], stay_non_gt_mask=cat["stay"] & ~cat["is_gt"], stay_loc_ids=cat["stay_loc"],
                      travel_mask=cat["travel"], prev_zone_idx=cat["prev"], dest_zone_idx=cat["dest"], progress_s=cat["prog"],
                      min_dt=min_dt)


# --------------------------------------------------------------------------------------------------------
# latent_ode collate
# --------------------------------------------------------------------------------------------------------
# ids fixed by the reference's feature tables (data_generator/feature_engineering.py:10-34): key order of
# MODE_FEATURES / PURPOSE_FEATURES
STAY_MODE_ID = 0         # MODE_ID_MAP["stay"]
TRAVEL_PURPOSE_ID = 5    # PURPOSE_ID_MAP["travel"]


def unify_and_interpolate_batch(batch: Sequence[Dict], train_on_interpolated_points: Optional[bool] = None,
                                purpose_groups=None) -> Dict:
    """Same 16-key dict as latent_ode/data_process/batching.py:109-128.  Each sample dict carries `times`, `trajectory_y`,
    `target_purpose_ids`, `target_mode_ids`, `target_purpose_features`, `target_mode_features`, `importance_weights`,
    `person_features`, `home_zone_features`, `work_zone_features`, `all_zone_features`, `num_zones`, `person_name` and
    (optionally) `config` with `train_on_interpolated_points` / `purpose_groups`."""
    cfg = batch[0].get("config", None)
    if train_on_interpolated_points is None:
        train_on_interpolated_points = bool(getattr(cfg, "train_on_interpolated_points", False))
    if purpose_groups is None:
        purpose_groups = getattr(cfg, "purpose_groups", None)
    device = batch[0]["person_features"].device
    B = len(batch)
    t_unified = torch.cat([s["times"] for s in batch]).unique(sorted=True)
    T = int(t_unified.numel())
    ar = torch.arange(T, device=device)

    S = max(int(s["times"].numel()) for s in batch)
    valid = torch.zeros((B, S), dtype=torch.bool, device=device)
    pos = torch.zeros((B, S), dtype=torch.long, device=device)
    for i, s in enumerate(batch):
        n = int(s["times"].numel())
        valid[i, :n] = True
        pos[i, :n] = torch.searchsorted(t_unified, s["times"].to(device))          # exact members of the union (:59-63)

    def scatter(key, fill, dtype, trailing=()):
        dense = torch.full((B, T) + tuple(trailing), fill, dtype=dtype, device=device)
        for i, s in enumerate(batch):
            n = int(s["times"].numel())
            dense[i, pos[i, :n]] = s[key].to(device=device, dtype=dtype)
        return dense

    pf = int(batch[0]["target_purpose_features"].shape[-1])
    mf = int(batch[0]["target_mode_features"].shape[-1])
    y_loc = scatter("trajectory_y", -1, torch.long)
    y_purp = scatter("target_purpose_ids", -1, torch.long)
    y_mode = scatter("target_mode_ids", -1, torch.long)
    y_purp_feat = scatter("target_purpose_features", 0.0, torch.float32, (pf,))
    y_mode_feat = scatter("target_mode_features", 0.0, torch.float32, (mf,))
    importance = scatter("importance_weights", 1.0, torch.float32)
    real = y_loc != -1
    loss_mask = torch.ones((B, T), device=device) if train_on_interpolated_points else real.to(torch.float32)
    # `loss_mask[i, idx] = 1` marks observation points even when their zone label is -1: use the scatter positions
    if not train_on_interpolated_points:
        loss_mask = torch.zeros((B, T), device=device)
        for i, s in enumerate(batch):
            loss_mask[i, pos[i, :int(s["times"].numel())]] = 1.0

    # previous / next real observation index for every grid point (:80-90), rows without observations stay 0
    has = real.any(dim=1, keepdim=True)
    first = torch.where(real, ar[None, :], torch.full((1, 1), T, dtype=torch.long, device=device)).amin(dim=1, keepdim=True)
    lastr = torch.where(real, ar[None, :], torch.full((1, 1), -1, dtype=torch.long, device=device)).amax(dim=1, keepdim=True)
    prev_strict = torch.where(real, ar[None, :], torch.full((1, 1), -1, dtype=torch.long, device=device))
    prev_strict = torch.cat([torch.full((B, 1), -1, dtype=torch.long, device=device), prev_strict[:, :-1]], dim=1).cummax(dim=1).values
    next_strict = torch.where(real, ar[None, :], torch.full((1, 1), T, dtype=torch.long, device=device))
    next_strict = torch.cat([next_strict[:, 1:], torch.full((B, 1), T, dtype=torch.long, device=device)], dim=1).flip(1).cummin(dim=1).values.flip(1)
    prev_real = torch.where(prev_strict >= 0, prev_strict, first)          # searchsorted(left) - 1, clamped to the first real index
    next_real = torch.where(next_strict < T, next_strict, lastr)            # searchsorted(right), clamped to the last real index
    zero = torch.zeros_like(prev_real)
    prev_real = torch.where(has, prev_real, zero)
    next_real = torch.where(has, next_real, zero)

    # fill purpose / mode ids of the interpolated points between two real observations (:93-105)
    a = torch.where(real, ar[None, :], torch.full((1, 1), -1, dtype=torch.long, device=device)).cummax(dim=1).values      # last real <= t
    b = torch.where(real, ar[None, :], torch.full((1, 1), T, dtype=torch.long, device=device)).flip(1).cummin(dim=1).values.flip(1)
    between = (~real) & (a >= 0) & (b < T)
    ac, bc = a.clamp_min(0), b.clamp_max(T - 1)
    sp, ep = y_purp.gather(1, ac), y_purp.gather(1, bc)
    sm, em = y_mode.gather(1, ac), y_mode.gather(1, bc)
    changed = sp != ep
    y_purp = torch.where(between, torch.where(changed, torch.full_like(sp, TRAVEL_PURPOSE_ID), sp), y_purp)
    trans_mode = torch.where(sm != STAY_MODE_ID, sm, em)
    y_mode = torch.where(between, torch.where(changed, trans_mode, sm), y_mode)

    return {
        "t_unified": t_unified,
        "y_loc_dense": y_loc,
        "y_purp_dense": y_purp,
        "y_mode_dense": y_mode,
        "y_purp_feat_dense": y_purp_feat,
        "y_mode_feat_dense": y_mode_feat,
        "loss_mask": loss_mask * importance,
        "prev_real_indices": prev_real,
        "next_real_indices": next_real,
        "person_features": torch.stack([s["person_features"] for s in batch]),
        "home_zone_features": torch.stack([s["home_zone_features"] for s in batch]),
        "work_zone_features": torch.stack([s["work_zone_features"] for s in batch]),
        "all_zone_features": batch[0]["all_zone_features"],
        "num_zones": batch[0]["num_zones"],
        "purpose_groups": purpose_groups,
        "person_names": [s["person_name"] for s in batch],
    }


# ------------------------------------------------------------------------------------------------------------------
# the layout the reference's own (stale) test module pins: /root/reference/test/test_data_batching.py:30-81
# ------------------------------------------------------------------------------------------------------------------
K_INTERNAL = 10     # test_data_batching.py:39-40: dense grid length (len(gt_union) - 1) * (K_INTERNAL - 1) + 1


def sde_collate_fn(samples: Sequence[Dict], k_internal: int = K_INTERNAL, time_match_tol: float = 1e-6) -> Dict:
    """Collate with the key names, shapes and invariants of the `sde_collate_fn` batch that
    /root/reference/test/test_data_batching.py checks (the function itself is no longer in the reference snapshot; the test
    module is the specification, SURVEY.md §4 invariants 1-5):

      gt_union_times [S_gt]  sorted union of every person's ground-truth times;  is_gt_union [B, S_gt]
      grid_times [S_dense]   every union gap cut into (k_internal - 1) equal sub-intervals, endpoints shared:
                             S_dense = (S_gt - 1) * (k_internal - 1) + 1;  is_gt_grid [S_dense] marks the union times, and
                             grid_times[is_gt_grid] == gt_union_times exactly                                  (:33-40)
      loc_emb_union [B, S_gt, E], purp_emb_union [B, S_gt, Ep] float32, anchor_union [B, S_gt]                   (:45-51)
                             a person's embedding on the union axis: held FLAT across a stay (the last snap's value), linearly
                             interpolated between origin and destination inside a travel segment                 (:53-68)
      segments_batch         ragged list over all persons' travel segments, dicts with keys exactly
                             {"b", "i0", "i1", "mode_id", "mode_proto"}; i0 < i1 are DENSE-grid indices of the segment's
                             end points, which are ground-truth grid points                                      (:70-81)

    `samples[b]` is a per-person dict: gt_times [S_b] (sorted), gt_loc_emb [S_b, E], gt_purp_emb [S_b, Ep], gt_anchor [S_b],
    segments = [{"t0", "t1", "mode_id", "mode_proto"}, ...] (travel legs; t0 / t1 are ground-truth times of that person).
    Tensorised over the union axis; person b of the input is row b of every output."""
    if k_internal < 2:
        raise ValueError("k_internal must be >= 2")
    dev = samples[0]["gt_times"].device
    B = len(samples)
    gt_union = torch.unique(torch.cat([s["gt_times"].to(torch.float32) for s in samples]), sorted=True)
    S = gt_union.numel()
    K1 = k_internal - 1
    # dense grid: exact union times at multiples of K1, equally spaced interior points in between
    if S > 1:
        frac = torch.arange(K1, device=dev, dtype=torch.float32) / K1                    # 0, 1/K1, ...
        lo, hi = gt_union[:-1, None], gt_union[1:, None]
        grid = torch.cat([(lo + (hi - lo) * frac[None, :]).reshape(-1), gt_union[-1:]])
        grid[::K1] = gt_union                                                            # bit-exact at the union points
    else:
        grid = gt_union.clone()
    is_gt_grid = torch.zeros(grid.numel(), dtype=torch.bool, device=dev)
    is_gt_grid[::K1] = True
    E = samples[0]["gt_loc_emb"].shape[-1]
    Ep = samples[0]["gt_purp_emb"].shape[-1]
    is_gt_union = torch.zeros((B, S), dtype=torch.bool, device=dev)
    loc = torch.zeros((B, S, E), dtype=torch.float32, device=dev)
    purp = torch.zeros((B, S, Ep), dtype=torch.float32, device=dev)
    anchor = torch.zeros((B, S), dtype=torch.float32, device=dev)
    segments_batch: List[Dict] = []
    for b, s in enumerate(samples):
        gt = s["gt_times"].to(torch.float32)
        # last ground-truth snap at or before every union time (clamped to the person's first snap)
        k = (torch.searchsorted(gt, gt_union + time_match_tol, right=True) - 1).clamp(0, gt.numel() - 1)
        hit = (gt[k] - gt_union).abs() <= time_match_tol
        is_gt_union[b] = hit
        le, pe = s["gt_loc_emb"].to(torch.float32)[k], s["gt_purp_emb"].to(torch.float32)[k]      # flat hold
        anchor[b] = torch.where(hit, s["gt_anchor"].to(torch.float32)[k], torch.zeros((), device=dev))
        for seg in s["segments"]:
            t0, t1 = float(seg["t0"]), float(seg["t1"])
            inside = (gt_union > t0 + time_match_tol) & (gt_union < t1 - time_match_tol)
            if bool(inside.any()):
                k0 = int(torch.searchsorted(gt, torch.tensor(t0 + time_match_tol, device=dev), right=True) - 1)
                k1 = int(torch.searchsorted(gt, torch.tensor(t1 - time_match_tol, device=dev), right=False))
                w = ((gt_union - t0) / (t1 - t0)).clamp(0, 1)[:, None]
                g0l, g1l = s["gt_loc_emb"].to(torch.float32)[k0], s["gt_loc_emb"].to(torch.float32)[k1]
                g0p, g1p = s["gt_purp_emb"].to(torch.float32)[k0], s["gt_purp_emb"].to(torch.float32)[k1]
                le = torch.where(inside[:, None], g0l + (g1l - g0l) * w, le)
                pe = torch.where(inside[:, None], g0p + (g1p - g0p) * w, pe)
            i0 = int(torch.argmin((gt_union - t0).abs())) * K1
            i1 = int(torch.argmin((gt_union - t1).abs())) * K1
            segments_batch.append({"b": b, "i0": i0, "i1": i1, "mode_id": seg["mode_id"], "mode_proto": seg["mode_proto"]})
        loc[b], purp[b] = le, pe
    return {"gt_union_times": gt_union, "grid_times": grid, "is_gt_grid": is_gt_grid, "is_gt_union": is_gt_union,
            "loc_emb_union": loc, "purp_emb_union": purp, "anchor_union": anchor, "segments_batch": segments_batch}
