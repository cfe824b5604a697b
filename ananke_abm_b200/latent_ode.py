"""`GenerativeODE` with the reference's constructor, parameter tree, forward signature and 8-tuple output, running the
ODE on the CUDA kernels (latent_ode is the reference's dopri5 call site and the `gnn_embed` slot).

Mirrors /root/reference/src/ananke_abm/models/latent_ode/architecture/model.py:9-17 (tanh ResidualBlock), :19-117
(ODEFunc, second-order branch with the potential-gradient correction) and :132-220 (GenerativeODE); config fields as in
latent_ode/config.py:18-71.  Same `state_dict` keys, so checkpoints are interchangeable.  The encoder / decoders /
einsum head are host-side PyTorch plumbing; every drift evaluation and all solver algebra inside `odeint` run in
libananke_b200.so (`ab200_drift_eval` with the closed-form correction term, `ab200_rk_combine_errnorm`).
The SDE branch (`enable_sde=True`, the reference default) runs forward-only (Euler-Maruyama sampling, `sdeint.py`); training
runs the ODE
branch and raises if asked for the SDE.  `zone_embed` may be a `gnn_embed.GATEmbed`: the slot a GAT fills (:171-173).
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
