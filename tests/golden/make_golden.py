"""Mint the golden vectors under tests/golden/ by running the UNMODIFIED reference modules.

Run in the authoring container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What it does
  * puts /root/reference/src on sys.path and imports the reference's own `ModeSepModel`,
    `GenerativeODE`, `build_union_batch`, `unify_and_interpolate_batch`, `load_csvs`,
    `build_person_and_shared`, `DataProcessor`, losses -- nothing re-typed;
  * installs `oracle.torchdiffeq_oracle` as `torchdiffeq` (the real package, pinned 0.2.5 in
    /root/reference/uv.lock:2896-2897, is not installable here) and a raising stub as `torchsde`;
  * config 1 of BASELINE.json: reference fixtures test/test_snaps_small.csv + test/test_periods_small.csv
    through the mode_sep loader, zones/dist_mat/persons from data_generator/generate_mock_csvs.main();
    seed 42 (mode_sep/config.py:12); forward + the training loss of mode_sep/train/train.py:101-159 +
    backward; everything frozen to mode_sep_fixture.npz;
  * the latent_ode mock batch (enable_sde=False, dopri5, the h0 noise drawn under seed 1234 and stored) ->
    latent_ode_fixture.npz, plus the collate dict of unify_and_interpolate_batch;
  * stand-alone RHS evaluations of both models on random states -> rhs_fixture.npz.
"""
from __future__ import annotations

import hashlib
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF / "src"))

from oracle import torchdiffeq_oracle as tdq  # noqa: E402

sys.modules["torchdiffeq"] = tdq
_sde = types.ModuleType("torchsde")


def _no_sde(*a, **k):
    raise RuntimeError("torchsde is not available; the SDE branch is out of scope")


_sde.sdeint = _no_sde
sys.modules["torchsde"] = _sde


def _np(x):
    return x.detach().cpu().numpy()


def mode_sep_fixture():
    from ananke_abm.data_generator import generate_mock_csvs
    from ananke_abm.models.mode_sep.config import ModeSepConfig
    from ananke_abm.models.mode_sep.data_process.data_paths import DataPaths
    from ananke_abm.models.mode_sep.data_process.io_csv import load_csvs
    from ananke_abm.models.mode_sep.data_process.data import build_person_and_shared
    from ananke_abm.models.mode_sep.data_process.batching import build_union_batch
    from ananke_abm.models.mode_sep.architecture.model import ModeSepModel
    from ananke_abm.models.mode_sep.architecture.losses import (total_loss, ce_at_snaps, mse_at_snaps,
                                                                 expected_distance_at_snaps)

    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            generate_mock_csvs.main()
        finally:
            os.chdir(cwd)
        data = Path(tmp) / "data"
        paths = DataPaths(snaps_csv=REF / "test/test_snaps_small.csv", periods_csv=REF / "test/test_periods_small.csv",
                          zones_csv=data / "zones.csv", dist_mat_csv=data / "dist_mat.csv",
                          persons_csv=data / "persons.csv")
        loaded = load_csvs(paths)
        zones_csv_text = (data / "zones.csv").read_text()
        persons_csv_text = (data / "persons.csv").read_text()
        dist_csv_text = (data / "dist_mat.csv").read_text()
    device = torch.device("cpu")
    cfg = ModeSepConfig()
    cfg.device = "cpu"
    persons, shared = build_person_and_shared(loaded, device)
    union = build_union_batch(persons, cfg, device)

    torch.manual_seed(cfg.seed)
    np.random.seed(cfg.seed)
    Z = loaded.id_maps.Z
    model = ModeSepModel(Z, cfg)
    model.train()

    home_idx = torch.tensor([p.home_zone_idx for p in persons], dtype=torch.long)
    work_idx = torch.tensor([p.work_zone_idx for p in persons], dtype=torch.long)
    traits = torch.stack([p.person_traits_raw for p in persons], dim=0)

    pred_emb, logits, v = model(union.times_union, home_idx, work_idx, traits)

    # y_path and y0 through the same code path the model used
    with torch.no_grad():
        p0 = model.class_table.detach()[home_idx]
        h = model._encode_context(traits, home_idx, work_idx)
        y0 = torch.cat([p0, torch.zeros_like(p0), h], dim=-1)
        y_path = tdq.odeint(model.odefunc, y0, union.times_union, method=cfg.ode_method, rtol=cfg.rtol, atol=cfg.atol)

    # training loss exactly as mode_sep/train/train.py:101-159
    B, T, _ = logits.shape
    y_union = torch.full((B, T), -1, dtype=torch.long)
    for i, p in enumerate(persons):
        sidx = union.snap_indices[i]
        m = sidx >= 0
        if m.any():
            y_union[i, m] = p.loc_ids[sidx[m]]
    base, parts = total_loss(config=cfg, logits=logits, pred_emb=pred_emb, y_union=y_union, is_gt_mask=union.is_gt_union,
                             dist_mat=shared.dist_mat, class_table=model.class_table, travel_mask=union.travel_mask,
                             prev_idx=union.prev_zone_idx, dest_idx=union.dest_zone_idx)
    y_stay, m_aux = union.stay_loc_ids, union.stay_non_gt_mask
    aux = cfg.w_stay_aux * (ce_at_snaps(logits, y_stay, m_aux) + mse_at_snaps(pred_emb, y_stay, model.class_table, m_aux)
                            + expected_distance_at_snaps(logits, y_stay, shared.dist_mat, m_aux))
    v_abs = v.norm(dim=-1)
    stay_vel = (v_abs[union.stay_non_gt_mask] ** 2).mean()
    v_m = v_abs[union.gt_interior_mask]
    move_vel = ((cfg.v_min_move - v_m).clamp(min=0.0) ** 2 + (v_m - cfg.v_max_move).clamp(min=0.0) ** 2).mean()
    total = base + aux + cfg.w_stay_vel_core * stay_vel + cfg.w_move_vel_hinge * move_vel
    total.backward()

    out = {
        "times_union": _np(union.times_union), "min_dt": np.float64(union.min_dt),
        "home_idx": _np(home_idx), "work_idx": _np(work_idx), "traits": _np(traits),
        "y_union": _np(y_union), "dist_mat": _np(shared.dist_mat),
        "y0": _np(y0), "y_path": _np(y_path), "pred_emb": _np(pred_emb), "logits": _np(logits), "v_t": _np(v),
        "labels": _np(logits.argmax(-1)), "loss_total": np.float64(total.item()),
        "loss_base": np.float64(base.item()),
        "csv_zones": np.array(zones_csv_text), "csv_persons": np.array(persons_csv_text),
        "csv_dist": np.array(dist_csv_text),
        "person_ids": np.array([p.person_id for p in persons]),
        "person_names": np.array([p.person_name for p in persons]),
    }
    for f in ("is_gt_union", "snap_indices", "stay_mask", "gt_interior_mask", "stay_non_gt_mask", "stay_loc_ids",
              "travel_mask", "prev_zone_idx", "dest_zone_idx", "progress_s"):
        out["ub_" + f] = _np(getattr(union, f))
    for k, p in model.state_dict().items():
        out["sd_" + k] = _np(p)
    for k, p in model.named_parameters():
        out["grad_" + k] = _np(p.grad) if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
    # per-person raw inputs so the GPU box can rebuild the union batch without the reference
    for i, p in enumerate(persons):
        out[f"p{i}_times_snap"] = _np(p.times_snap)
        out[f"p{i}_loc_ids"] = _np(p.loc_ids)
        out[f"p{i}_stay_segments"] = np.array(p.stay_segments, dtype=np.float64).reshape(-1, 3)
    np.savez_compressed(HERE / "mode_sep_fixture.npz", **out)
    sha = hashlib.sha256(_np(union.times_union).tobytes()).hexdigest()[:16]
    print("mode_sep fixture: B=%d T=%d Z=%d  times sha256[:16]=%s  loss=%.6f" % (B, T, Z, sha, total.item()))


def latent_fixture():
    from torch.utils.data import DataLoader
    from ananke_abm.models.latent_ode.config import GenerativeODEConfig
    from ananke_abm.models.latent_ode.data_process.data import DataProcessor, LatentODEDataset
    from ananke_abm.models.latent_ode.data_process.batching import unify_and_interpolate_batch
    from ananke_abm.models.latent_ode.architecture.model import GenerativeODE

    device = torch.device("cpu")
    cfg = GenerativeODEConfig()
    cfg.enable_sde = False
    proc = DataProcessor(device, cfg)
    ds = LatentODEDataset([1, 2], proc)
    loader = DataLoader(ds, batch_size=2, collate_fn=unify_and_interpolate_batch)
    batch = next(iter(loader))

    torch.manual_seed(42)
    model = GenerativeODE(person_feat_dim=batch["person_features"].shape[-1],
                          num_zone_features=batch["all_zone_features"].shape[-1], config=cfg)
    args = (batch["person_features"], batch["home_zone_features"], batch["work_zone_features"],
            batch["y_purp_feat_dense"][:, 0], batch["y_mode_feat_dense"][:, 0], batch["t_unified"],
            batch["all_zone_features"])
    torch.manual_seed(1234)
    eps = torch.randn(batch["person_features"].shape[0], cfg.hidden_dim)     # the draw randn_like will make
    torch.manual_seed(1234)
    outs = model(*args)
    solver = tdq._LAST_SOLVER["solver"]
    names = ["loc_logits", "loc_embed", "purp_logits", "mode_logits", "purp_feat", "mode_feat", "h0_mu", "h0_log_var"]
    out = {n: _np(o) for n, o in zip(names, outs)}
    out["eps"] = _np(eps)
    out["n_accepted"] = np.int64(solver.n_accepted)
    out["n_rejected"] = np.int64(solver.n_rejected)
    out["step_log"] = np.array([(a, b, float(c)) for a, b, c in solver.step_log], dtype=np.float64)
    # a loss with a simple closed form so the gradient check does not depend on the 8-term composite loss
    loss = outs[0].pow(2).mean() + outs[2].pow(2).mean() + outs[3].pow(2).mean() + outs[1].pow(2).mean()
    loss.backward()
    out["loss"] = np.float64(loss.item())
    for k, v in batch.items():
        if torch.is_tensor(v):
            out["batch_" + k] = _np(v)
    out["batch_person_names"] = np.array(batch["person_names"])
    out["batch_purpose_groups"] = np.array(list(batch["purpose_groups"]))
    out["batch_num_zones"] = np.int64(batch["num_zones"])
    for k, p in model.state_dict().items():
        out["sd_" + k] = _np(p)
    for k, p in model.named_parameters():
        out["grad_" + k] = _np(p.grad) if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
    # raw per-person samples, so the collate function can be re-run without the reference
    for i, pid in enumerate([1, 2]):
        s = proc.get_data(pid)
        for k, v in s.items():
            if torch.is_tensor(v):
                out[f"s{i}_{k}"] = _np(v)
        out[f"s{i}_person_name"] = np.array(s["person_name"])
    np.savez_compressed(HERE / "latent_ode_fixture.npz", **out)
    print("latent fixture: T=%d accepted=%d rejected=%d loss=%.6f" % (len(batch["t_unified"]), solver.n_accepted,
                                                                      solver.n_rejected, loss.item()))


def latent_loss_fixture():
    """The reference's own `calculate_composite_loss` (latent_ode/architecture/loss.py:7-100, unmodified) on the frozen latent
    fixture: the model outputs of latent_ode_fixture.npz as leaves -> the nine returned values and the gradients with respect to
    every model output and to the zone encoder -> latent_loss_fixture.npz."""
    from ananke_abm.models.latent_ode.config import GenerativeODEConfig
    from ananke_abm.models.latent_ode.data_process.data import DataProcessor
    from ananke_abm.models.latent_ode.architecture.model import GenerativeODE
    from ananke_abm.models.latent_ode.architecture.loss import calculate_composite_loss
    g = np.load(HERE / "latent_ode_fixture.npz", allow_pickle=False)
    cfg = GenerativeODEConfig()
    cfg.enable_sde = False
    proc = DataProcessor(torch.device("cpu"), cfg)
    model = GenerativeODE(person_feat_dim=g["batch_person_features"].shape[-1], num_zone_features=g["batch_all_zone_features"].shape[-1],
                          config=cfg)
    model.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd_")})
    names = ["loc_logits", "loc_embed", "purp_logits", "mode_logits", "purp_feat", "mode_feat", "h0_mu", "h0_log_var"]
    outs = [torch.from_numpy(g[n]).clone().requires_grad_(True) for n in names]
    batch = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("batch_") and g[k].dtype.kind in "fiub" and g[k].ndim > 0}
    vals = calculate_composite_loss(batch, tuple(outs), model, proc.distance_matrix, cfg)
    vals[0].backward()
    out = {"distance_matrix": _np(proc.distance_matrix)}
    for n, v in zip(["total", "classification", "embedding", "distance", "purpose_class", "purpose_mse", "mode_class", "mode_mse", "kl"], vals):
        out["loss_" + n] = np.float64(v.item())
    for n, o in zip(names, outs):
        out["grad_" + n] = _np(o.grad) if o.grad is not None else np.zeros(tuple(o.shape), np.float32)
    out["grad_zone_feature_encoder.weight"] = _np(model.zone_feature_encoder.weight.grad)
    out["grad_zone_feature_encoder.bias"] = _np(model.zone_feature_encoder.bias.grad)
    for k in ("loss_weight_classification", "loss_weight_embedding", "loss_weight_distance", "loss_weight_purpose_class",
              "loss_weight_mode_class", "loss_weight_purpose_mse", "loss_weight_mode_mse", "kl_weight"):
        out["cfg_" + k] = np.float64(getattr(cfg, k))
    np.savez_compressed(HERE / "latent_loss_fixture.npz", **out)
    print("latent loss fixture written", {k: float(v) for k, v in out.items() if k.startswith("loss_")})


def rhs_fixture():
    from ananke_abm.models.mode_sep.config import ModeSepConfig
    from ananke_abm.models.mode_sep.architecture.model import ModeSepModel
    from ananke_abm.models.latent_ode.config import GenerativeODEConfig
    from ananke_abm.models.latent_ode.architecture.model import GenerativeODE

    out = {}
    torch.manual_seed(42)
    m = ModeSepModel(8, ModeSepConfig())
    g = torch.Generator().manual_seed(7)
    y = torch.randn(33, 160, generator=g) * 0.5
    for i, tv in enumerate([0.0, 7.25, 23.5]):
        out[f"ms_t{i}"] = np.float32(tv)
        out[f"ms_f{i}"] = _np(m.odefunc(torch.tensor(tv), y))
    out["ms_y"] = _np(y)
    for k, p in m.state_dict().items():
        out["ms_sd_" + k] = _np(p)

    cfg = GenerativeODEConfig()
    cfg.enable_sde = False
    torch.manual_seed(42)
    gm = GenerativeODE(8, 7, cfg)
    y2 = torch.randn(33, 64, generator=g) * 0.7
    for i, tv in enumerate([0.0, 7.25, 23.5]):
        out[f"lo_t{i}"] = np.float32(tv)
        out[f"lo_f{i}"] = _np(gm.ode_func(torch.tensor(tv), y2.clone()))
    out["lo_y"] = _np(y2)
    for k, p in gm.state_dict().items():
        out["lo_sd_" + k] = _np(p)
    np.savez_compressed(HERE / "rhs_fixture.npz", **out)
    print("rhs fixture written")


def loss_terms_fixture():
    """The reference's own loss functions (mode_sep/architecture/losses.py:14-44, unmodified) on the frozen fixture
    logits: the terms the fused head reproduces from (pred_emb, class_table) -> loss_terms_fixture.npz."""
    from ananke_abm.models.mode_sep.architecture.losses import ce_at_snaps, mse_at_snaps, expected_distance_at_snaps
    g = np.load(HERE / "mode_sep_fixture.npz")
    logits = torch.from_numpy(g["logits"])
    pred_emb = torch.from_numpy(g["pred_emb"])
    table = torch.from_numpy(g["sd_class_table"])
    dist = torch.from_numpy(g["dist_mat"])
    y_union, is_gt = torch.from_numpy(g["y_union"]), torch.from_numpy(g["ub_is_gt_union"])
    y_stay, m_aux = torch.from_numpy(g["ub_stay_loc_ids"]), torch.from_numpy(g["ub_stay_non_gt_mask"])
    out = {
        "ce_gt": np.float64(ce_at_snaps(logits, y_union, is_gt).item()),
        "dist_gt": np.float64(expected_distance_at_snaps(logits, y_union, dist, is_gt).item()),
        "mse_gt": np.float64(mse_at_snaps(pred_emb, y_union, table, is_gt).item()),
        "ce_aux": np.float64(ce_at_snaps(logits, y_stay, m_aux).item()),
        "dist_aux": np.float64(expected_distance_at_snaps(logits, y_stay, dist, m_aux).item()),
        "n_gt": np.int64(int(is_gt.sum())), "n_aux": np.int64(int(m_aux.sum())),
    }
    np.savez_compressed(HERE / "loss_terms_fixture.npz", **out)
    print("loss terms fixture written", {k: float(v) for k, v in out.items()})


if __name__ == "__main__":
    assert REF.exists(), "the reference tree is only present in the authoring container"
    if len(sys.argv) > 1 and sys.argv[1] == "latent_loss":  # only the latent composite loss (reads the frozen latent fixture)
        latent_loss_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "losses":       # only the loss terms (reads the frozen mode_sep fixture)
        loss_terms_fixture()
        sys.exit(0)
    mode_sep_fixture()
    latent_fixture()
    rhs_fixture()
    loss_terms_fixture()
    latent_loss_fixture()
