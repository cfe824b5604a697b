"""Restatement (test infrastructure) of the mode_sep training loss so that GPU tests can reproduce the golden
gradients without /root/reference.  Follows /root/reference/src/ananke_abm/models/mode_sep/architecture/losses.py:14-158
and mode_sep/train/train.py:101-159 term by term; weights from mode_sep/config.py:41-60."""
import torch
import torch.nn.functional as F

W = dict(w_ce=1.0, w_mse=0.5, w_dist=0.5, w_stay_aux=0.9, w_stay_vel_core=5.0, w_move_vel_hinge=1.0, v_min_move=0.2,
         v_max_move=1.0, w_travel_margin=1.0, w_travel_mono=0.5, m_travel=0.10, epsilon_mono=0.01)


def _ce(logits, y, mask):
    return F.cross_entropy(logits[mask], y[mask], reduction="mean") if mask.any() else logits.new_zeros(())


def _mse(pred, y, table, mask):
    if not mask.any():
        return pred.new_zeros(())
    return (pred - table[y.clamp(min=0)]).pow(2).sum(-1)[mask].mean()


def _edist(logits, y, dist, mask):
    if not mask.any():
        return logits.new_zeros(())
    return (dist[y.clamp(min=0)] * torch.softmax(logits, -1)).sum(-1)[mask].mean()


def _d2c(pred, table, idx):
    return (pred - table[idx.clamp_min(0)]).pow(2).sum(-1).sqrt()


def mode_sep_training_loss(g, pred, logits, v, class_table, dev):
    tt = lambda k: torch.from_numpy(g[k]).to(dev)   # noqa: E731
    y_union, is_gt, dist = tt("y_union"), tt("ub_is_gt_union"), tt("dist_mat")
    travel, prev, dest = tt("ub_travel_mask"), tt("ub_prev_zone_idx"), tt("ub_dest_zone_idx")
    base = W["w_ce"] * _ce(logits, y_union, is_gt) + W["w_mse"] * _mse(pred, y_union, class_table, is_gt) \
        + W["w_dist"] * _edist(logits, y_union, dist, is_gt)
    if travel.any():
        dp, dd = _d2c(pred, class_table, prev), _d2c(pred, class_table, dest)
        base = base + W["w_travel_margin"] * (W["m_travel"] - (dp - dd))[travel].clamp(min=0.0).mean()
        pair = travel[:, :-1] & travel[:, 1:] & (prev[:, :-1] == prev[:, 1:]) & (dest[:, :-1] == dest[:, 1:])
        if pair.any():
            away = (dp[:, :-1][pair] - dp[:, 1:][pair] + W["epsilon_mono"]).clamp(min=0.0)
            toward = (dd[:, 1:][pair] - dd[:, :-1][pair] + W["epsilon_mono"]).clamp(min=0.0)
            base = base + W["w_travel_mono"] * (away.mean() + toward.mean()) * 0.5
    y_stay, m_aux = tt("ub_stay_loc_ids"), tt("ub_stay_non_gt_mask")
    aux = W["w_stay_aux"] * (_ce(logits, y_stay, m_aux) + _mse(pred, y_stay, class_table, m_aux) + _edist(logits, y_stay, dist, m_aux))
    v_abs = v.norm(dim=-1)
    stay_vel = (v_abs[m_aux] ** 2).mean()
    v_m = v_abs[tt("ub_gt_interior_mask")]
    move_vel = ((W["v_min_move"] - v_m).clamp(min=0.0) ** 2 + (v_m - W["v_max_move"]).clamp(min=0.0) ** 2).mean()
    return base + aux + W["w_stay_vel_core"] * stay_vel + W["w_move_vel_hinge"] * move_vel
