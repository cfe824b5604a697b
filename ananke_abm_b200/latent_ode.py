"""`GenerativeODE` with the reference's constructor, parameter tree, forward signature and 8-tuple output, running the
ODE on the CUDA kernels (latent_ode is the reference's dopri5 call site and the `gnn_embed` slot).

Mirrors /root/reference/src/ananke_abm/models/latent_ode/architecture/model.py:9-17 (tanh ResidualBlock), :19-117
(ODEFunc, second-order branch with the potential-gradient correction) and :132-220 (GenerativeODE); config fields as in
latent_ode/config.py:18-71.  Same `state_dict` keys, so checkpoints are interchangeable.  The encoder / decoders /
einsum head are host-side PyTorch plumbing; every drift evaluation and all solver algebra inside `odeint` run in
libananke_b200.so (`ab200_drift_eval` with the closed-form correction term, `ab200_rk_combine_errnorm`).
Training back-propagates through dopri5 on kernels (every drift evaluation carries `ab200_drift_vjp` as its backward); the SDE
branch (`enable_sde=True`, the reference default) samples with Euler-Maruyama (`sdeint.py`) and trains through the sampler the
same way.  `calculate_composite_loss` is the 8-term loss of latent_ode/architecture/loss.py.  `zone_embed` may be a
`gnn_embed.GATEmbed`: the slot a GAT fills (:171-173).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch
from torch import nn

from .drift import _ResidualBlock
from .odeint import odeint

PURPOSE_GROUPS = ("home", "work", "education", "shopping", "social", "travel")     # feature_engineering.py:24-33


@dataclass
class GenerativeODEConfig:                  # latent_ode/config.py:18-71 (fields the model reads, same defaults)
    hidden_dim: int = 32
    encoder_hidden_dim: int = 64
    ode_hidden_dim: int = 128
    zone_embed_dim: int = 8
    purpose_feature_dim: int = 4
    mode_feature_dim: int = 4
    num_residual_blocks: int = 2
    correction_strength: float = 1.0
    use_second_order_sde: bool = True
    train_on_interpolated_points: bool = False
    ode_method: str = "dopri5"
    enable_sde: bool = False                # reference default True; only the ODE branch exists here
    sde_noise_strength: float = 0.1
    num_modes: int = 5
    purpose_groups: tuple = field(default_factory=lambda: PURPOSE_GROUPS)
    # loss weights (latent_ode/config.py:38-48), read by `calculate_composite_loss`
    kl_weight: float = 0.5
    loss_weight_classification: float = 1.0
    loss_weight_embedding: float = 0.5
    loss_weight_distance: float = 2.0
    loss_weight_purpose_class: float = 0.75
    loss_weight_mode_class: float = 1.0
    loss_weight_purpose_mse: float = 0.5
    loss_weight_mode_mse: float = 0.5


class ODEFunc(nn.Module):                   # model.py:19-117 -- parameter holder; `describe_drift` recognises this shape
    _ab200_kernel_only = True     # forward() is a kernel call (ab200_drift_eval, with ab200_drift_vjp as its autograd backward)
    def __init__(self, config, state_dim: int, position_dim: int, hidden_dim: int, num_residual_blocks: int):
        super().__init__()
        self.config, self.state_dim, self.position_dim = config, state_dim, position_dim
        self.IS_MOVING_DIM = 0
        self.IS_STATIONARY_DIM = 0
        if not config.use_second_order_sde:
            raise NotImplementedError("only the second-order drift is instantiated in the CUDA library")
        layers = [nn.Linear(state_dim + config.hidden_dim + 2, hidden_dim), nn.ReLU()]
        layers += [_ResidualBlock(hidden_dim, "tanh") for _ in range(num_residual_blocks)]
        layers.append(nn.Linear(hidden_dim, position_dim))
        self.net = nn.Sequential(*layers)

    def forward(self, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        from .drift import describe_drift
        from .odeint import drift_apply
        spec = describe_drift(self)
        if spec is None:
            raise RuntimeError("drift shape not instantiated in libananke_b200.so")
        return drift_apply(spec, t, y)

    # torchsde interface of the reference (model.py:119-130): noise on the state only, none on the context h
    noise_type, sde_type = "diagonal", "ito"

    def f(self, t, y):
        return self.forward(t, y)

    def g(self, t, y):
        noise = y.new_zeros(y.shape)
        noise[:, : self.state_dim] = self.config.sde_noise_strength
        return noise


class GenerativeODE(nn.Module):
    def __init__(self, person_feat_dim: int, num_zone_features: int, config: GenerativeODEConfig,
                 zone_embed: Optional[nn.Module] = None):
        super().__init__()
        self.config = config
        self.zone_feature_encoder = zone_embed if zone_embed is not None else nn.Linear(num_zone_features, config.zone_embed_dim)
        enc_in = person_feat_dim + 2 * config.zone_embed_dim + config.purpose_feature_dim + config.mode_feature_dim
        self.encoder = nn.Sequential(nn.Linear(enc_in, config.encoder_hidden_dim), nn.ReLU(),
                                     nn.Linear(config.encoder_hidden_dim, 2 * config.hidden_dim))
        self.position_dim = config.zone_embed_dim + config.purpose_feature_dim + config.mode_feature_dim
        self.state_dim = 2 * self.position_dim if config.use_second_order_sde else self.position_dim
        self.ode_func = ODEFunc(config, self.state_dim, self.position_dim, config.ode_hidden_dim, config.num_residual_blocks)
        self.decoder_loc = nn.Linear(config.zone_embed_dim, config.zone_embed_dim)
        self.decoder_purpose = nn.Linear(config.purpose_feature_dim, len(config.purpose_groups))
        self.decoder_mode = nn.Linear(config.mode_feature_dim, config.num_modes)

    def forward(self, person_features, home_zone_features, work_zone_features, initial_purpose_features, initial_mode_features,
                times, all_zone_features, eps: Optional[torch.Tensor] = None, **odeint_kwargs) -> Tuple[torch.Tensor, ...]:
        """`eps` fixes the h0 reparameterisation noise (model.py:181 draws `randn_like`); extra keyword arguments
        (rtol, atol) go to `odeint` -- the reference uses torchdiffeq's defaults."""
        cfg = self.config
        cand = self.zone_feature_encoder(all_zone_features)
        home = self.zone_feature_encoder(home_zone_features)
        work = self.zone_feature_encoder(work_zone_features)
        enc_in = torch.cat([person_features, home, work, initial_purpose_features, initial_mode_features], dim=-1)
        h0_mu, h0_log_var = self.encoder(enc_in).split(cfg.hidden_dim, dim=-1)
        noise = torch.randn_like(h0_mu) if eps is None else eps
        h0 = h0_mu + torch.exp(0.5 * h0_log_var) * noise
        p0 = torch.cat([home, initial_purpose_features, initial_mode_features], dim=-1)
        s0 = torch.cat([p0, torch.zeros_like(p0)], dim=-1)
        y0 = torch.cat([s0, h0], dim=-1)
        if cfg.enable_sde:
            # model.py:192-194: sdeint(self.ode_func, y0, times, method='euler', dt=0.01); differentiable (the reference's
            # default training path goes through it): every Euler-Maruyama step carries its own backward
            from .sdeint import sdeint
            path = sdeint(self.ode_func, y0, times, method="euler", dt=0.01, options={"dtype": torch.float32},
                          seed=odeint_kwargs.pop("seed", None))
        else:
            path = odeint(self.ode_func, y0, times, method=cfg.ode_method, options={"dtype": torch.float32}, **odeint_kwargs)
        pred_y = path.permute(1, 0, 2)
        pred_s, _ = torch.split(pred_y, [self.state_dim, cfg.hidden_dim], dim=-1)
        pred_p = torch.split(pred_s, self.position_dim, dim=-1)[0]
        loc_embed, purp_feat, mode_feat = torch.split(pred_p, [cfg.zone_embed_dim, cfg.purpose_feature_dim, cfg.mode_feature_dim], dim=-1)
        loc_logits = torch.einsum("bsd,zd->bsz", self.decoder_loc(loc_embed), cand)
        return (loc_logits, loc_embed, self.decoder_purpose(purp_feat), self.decoder_mode(mode_feat), purp_feat, mode_feat,
                h0_mu, h0_log_var)


def calculate_composite_loss(batch, model_outputs, model, distance_matrix, config):
    """The latent model's 8-term training loss with the reference's signature and return tuple
    (latent_ode/architecture/loss.py:7-100; called at latent_ode/train/train.py:67-70):

        (total, classification, embedding, distance, purpose_class, purpose_mse, mode_class, mode_mse, kl)

    `batch` is the collate dict of `unify_and_interpolate_batch`, `model_outputs` the 8-tuple of `GenerativeODE.forward`.
    Every term is a masked mean over the [B, T] grid weighted by `loss_mask`; the targets of the embedding term are the zone
    embeddings of the previous / next REAL observation blended by elapsed time.  All inputs are [B, T, <= 8]-sized at the
    reference's scale (B = 2, Z = 8): this is host-side glue in PyTorch ops on the tensors' device, term for term the
    reference's arithmetic; the ODE solve that produces `model_outputs` is what runs in the CUDA library."""
    import torch.nn.functional as F
    (loc_logits, loc_embed, purp_logits, mode_logits, purp_feat, mode_feat, mu, log_var) = model_outputs
    t = batch["t_unified"]
    y_loc, y_purp, y_mode = batch["y_loc_dense"], batch["y_purp_dense"], batch["y_mode_dense"]
    mask = batch["loss_mask"]
    msum = mask.sum()
    B = loc_logits.shape[0]

    def masked_ce(logits, target):
        ce = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), target.reshape(-1), ignore_index=-1, reduction="none")
        return (ce * mask.reshape(-1)).sum() / msum

    def masked_mse(pred, target):
        return (F.mse_loss(pred, target, reduction="none").mean(dim=-1) * mask).sum() / msum

    loss_classification = masked_ce(loc_logits, y_loc)
    cand = model.zone_feature_encoder(batch["all_zone_features"])
    prev_ids = torch.gather(y_loc, 1, batch["prev_real_indices"])
    next_ids = torch.gather(y_loc, 1, batch["next_real_indices"])
    prev_e, next_e = cand[prev_ids.clamp(min=0)], cand[next_ids.clamp(min=0)]
    t_prev, t_next = t[batch["prev_real_indices"]], t[batch["next_real_indices"]]
    w_next = torch.clamp((t.unsqueeze(0) - t_prev) / (t_next - t_prev + 1e-8), 0, 1).unsqueeze(-1)
    loss_embedding = masked_mse(loc_embed, (1 - w_next) * prev_e + w_next * next_e)
    pred_ids = torch.argmax(loc_logits, dim=2)
    loss_distance = (distance_matrix[pred_ids, y_loc.clamp(min=0)] * mask).sum() / msum
    loss_purpose_class = masked_ce(purp_logits, y_purp)
    loss_purpose_mse = masked_mse(purp_feat, batch["y_purp_feat_dense"])
    loss_mode_class = masked_ce(mode_logits, y_mode)
    loss_mode_mse = masked_mse(mode_feat, batch["y_mode_feat_dense"])
    kl = -0.5 * torch.sum(1 + log_var - mu.pow(2) - log_var.exp()) / B
    total = (config.loss_weight_classification * loss_classification + config.loss_weight_embedding * loss_embedding
             + config.loss_weight_distance * loss_distance + config.loss_weight_purpose_class * loss_purpose_class
             + config.loss_weight_mode_class * loss_mode_class + config.loss_weight_purpose_mse * loss_purpose_mse
             + config.loss_weight_mode_mse * loss_mode_mse + config.kl_weight * kl)
    return (total, loss_classification, loss_embedding, loss_distance, loss_purpose_class, loss_purpose_mse, loss_mode_class,
            loss_mode_mse, kl)
