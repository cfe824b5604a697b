"""`ModeSepModel` with the reference's constructor, parameter tree and forward signature, running the ODE
on the fused CUDA kernels.

Mirrors /root/reference/src/ananke_abm/models/mode_sep/architecture/model.py:92-201 (and config.py:9-71):
same `state_dict` keys (checkpoints are interchangeable), same inputs
`forward(times_union[T], home_idx[B], work_idx[B], person_traits_raw[B,2])`, same outputs
`(pred_emb[B,T,E], logits[B,T,Z], v_t[B,T,E])`.  The pre/post glue (embedding gathers, the 18->128->32 context
encoder, the decoder and cosine head) is host-side PyTorch plumbing on the GPU; the drift net and the
Runge-Kutta integration -- every FLOP inside `odeint` -- run in libananke_b200.so.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch
from torch import nn

from .drift import _ResidualBlock
from .odeint import odeint, odeint_adjoint


@dataclass
class ModeSepConfig:                       # mode_sep/config.py:9-71 (fields the model reads, same defaults)
    seed: int = 42
    device: str = "cuda"
    emb_dim: int = 64
    context_dim: int = 32
    zone_emb_dim: int = 8
    hidden_dim: int = 128
    num_res_blocks: int = 2
    K_internal: int = 8
    ode_method: str = "rk4"
    rtol: float = 1e-5
    atol: float = 1e-5
    time_match_tol: float = 1e-6
    enable_sde: bool = False
    sde_noise_strength: float = 0.01
    sde_method: str = "euler"              # config.py:34-35
    sde_dt: float = 0.01
    softmax_tau: float = 0.2
    # loss weights and thresholds (config.py:38-53), read by losses.mode_sep_total_loss
    w_ce: float = 1.0
    w_mse: float = 0.5
    w_dist: float = 0.5
    w_stay_aux: float = 0.9
    w_stay_vel_core: float = 5.0
    w_move_vel_hinge: float = 1.0
    v_min_move: float = 0.2
    v_max_move: float = 1.0
    w_travel_margin: float = 1.0
    w_travel_mono: float = 0.5
    m_travel: float = 0.10
    epsilon_mono: float = 0.01
    lr: float = 1e-3
    weight_decay: float = 0.0
    grad_clip: float = 1.0
    precision: str = "f32"                 # ananke_b200 extension: 'f32' | 'bf16' (tensor-core drift GEMMs)
    error_norm: str = "shard"              # dopri5 across ranks: 'global' = one RMS norm over all shards (single-process parity)


def _solver_options(config) -> dict:
    opt = {"precision": getattr(config, "precision", "f32")}
    if getattr(config, "adjoint", False) and getattr(config, "adjoint_mode", None):
        opt["adjoint_mode"] = config.adjoint_mode          # "continuous" (torchdiffeq semantics, default) | "discrete"
    if getattr(config, "adjoint", False) and getattr(config, "adjoint_fused", None) is not None:
        opt["adjoint_fused"] = config.adjoint_fused      # launch structure of the tensor-core continuous adjoint (adjoint_tc.py)
    if config.ode_method == "rk4" and getattr(config, "step_size", None):
        opt["step_size"] = float(config.step_size)          # torchdiffeq fixed-grid option: solver grid t[0] + k step_size
    if config.ode_method == "dopri5" and opt["precision"] == "bf16":
        opt["error_norm"] = getattr(config, "error_norm", "shard")
        if getattr(config, "forward_operands", None):       # "fp16x2" (default of the solver) | "fp16" | "bf16"
            opt["forward_operands"] = config.forward_operands
        if getattr(config, "saved_operands", None):         # "all" (default of the solver) | "inputs" | "none": memory vs backward traffic
            opt["saved_operands"] = config.saved_operands
    return opt


class ODEFunc(nn.Module):                  # model.py:30-38 -- parameter holder; evaluated through WrappedSDE
    def __init__(self, emb_dim: int, context_dim: int, hidden_dim: int, num_blocks: int):
        super().__init__()
        layers = [nn.Linear(2 * emb_dim + context_dim + 2, hidden_dim), nn.ReLU()]
        layers += [_ResidualBlock(hidden_dim, "relu") for _ in range(num_blocks)]
        layers.append(nn.Linear(hidden_dim, emb_dim))
        self.net = nn.Sequential(*layers)


class WrappedSDE(nn.Module):               # model.py:49-73 -- f(t, y) = [v, net([p, v, h, sin, cos]), 0]
    _ab200_kernel_only = True     # forward() is a kernel call (ab200_drift_eval, with ab200_drift_vjp as its autograd backward)
    def __init__(self, func: ODEFunc, emb_dim: int, context_dim: int):
        super().__init__()
        self.func, self.emb_dim, self.context_dim = func, emb_dim, context_dim

    def forward(self, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        from .drift import describe_drift
        from .odeint import drift_apply
        spec = describe_drift(self)
        if spec is None:
            raise RuntimeError("drift shape not instantiated in libananke_b200.so")
        return drift_apply(spec, t, y)

    # torchsde interface of the reference (model.py:75-89): drift f = forward, unit diagonal diffusion on [p, v] only
    def f(self, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return self.forward(t, y)

    def g(self, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        noise = y.new_zeros(y.shape)
        noise[:, : 2 * self.emb_dim] = 1.0
        return noise


class _ScaledSDE(nn.Module):               # model.py:160-174: diffusion scaled by config.sde_noise_strength
    noise_type, sde_type = "diagonal", "ito"

    def __init__(self, base: WrappedSDE, scale: float):
        super().__init__()
        self.base, self.scale = base, scale

    def f(self, t, y):
        return self.base.f(t, y)

    def g(self, t, y):
        return self.base.g(t, y) * self.scale


class ModeSepModel(nn.Module):
    def __init__(self, Z: int, config: ModeSepConfig):
        super().__init__()
        self.config, self.Z = config, Z
        E, H = config.emb_dim, config.context_dim
        self.class_table = nn.Parameter(torch.empty(Z, E))
        nn.init.xavier_uniform_(self.class_table)
        self.zone_embed = nn.Embedding(Z, config.zone_emb_dim)
        self.context_encoder = nn.Sequential(nn.Linear(2 + 2 * config.zone_emb_dim, config.hidden_dim), nn.ReLU(),
                                             nn.Linear(config.hidden_dim, H))
        self.odefunc = WrappedSDE(ODEFunc(E, H, config.hidden_dim, config.num_res_blocks), E, H)
        self.decoder = nn.Sequential(nn.Linear(E, config.hidden_dim), nn.ReLU(), nn.Linear(config.hidden_dim, E))

    def _encode_context(self, traits_raw, home_idx, work_idx):
        raw = torch.cat([traits_raw, self.zone_embed(home_idx), self.zone_embed(work_idx)], dim=-1)
        return self.context_encoder(raw)

    def initial_state(self, home_idx, work_idx, person_traits_raw) -> torch.Tensor:
        p0 = self.class_table.detach()[home_idx]
        h = self._encode_context(person_traits_raw, home_idx, work_idx)
        return torch.cat([p0, torch.zeros_like(p0), h], dim=-1)

    def integrate(self, y0: torch.Tensor, times_union: torch.Tensor) -> torch.Tensor:
        if self.config.enable_sde and self.config.sde_noise_strength > 0.0:
            # model.py:158-182: Euler-Maruyama sampling (forward only here; sdeint refuses calls that need gradients)
            from .sdeint import sdeint
            return sdeint(_ScaledSDE(self.odefunc, self.config.sde_noise_strength), y0, times_union,
                          method=self.config.sde_method, dt=self.config.sde_dt, seed=getattr(self.config, "sde_seed", None))
        # config.adjoint (not a reference field): go through the odeint_adjoint seam (latent_ode/architecture/ode_components.py:50)
        if getattr(self.config, "adjoint", False):
            extra = {}
            if getattr(self.config, "adjoint_options", None) and getattr(self.config, "adjoint_mode", "continuous") != "discrete":
                extra["adjoint_options"] = dict(self.config.adjoint_options)      # e.g. {"norm": "seminorm"} (torchdiffeq adjoint.py)
            return odeint_adjoint(self.odefunc, y0, times_union, method=self.config.ode_method, rtol=self.config.rtol,
                                  atol=self.config.atol, options=_solver_options(self.config), **extra)
        return odeint(self.odefunc, y0, times_union, method=self.config.ode_method, rtol=self.config.rtol,
                      atol=self.config.atol, options=_solver_options(self.config))

    def head(self, y_path: torch.Tensor):
        E, H = self.config.emb_dim, self.config.context_dim
        yb = y_path.permute(1, 0, 2)
        p_t, v_t, _ = torch.split(yb, [E, E, H], dim=-1)
        pred_emb = self.decoder(p_t)
        table_norm = self.class_table / (self.class_table.norm(dim=-1, keepdim=True) + 1e-8)
        emb_norm = pred_emb / (pred_emb.norm(dim=-1, keepdim=True) + 1e-8)
        logits = torch.einsum("bte,ze->btz", emb_norm, table_norm) / self.config.softmax_tau
        return pred_emb, logits, v_t

    def training_loss(self, times_union, home_idx, work_idx, person_traits_raw, union, y_union, dist_mat):
        """One training forward of mode_sep/train/train.py:94-159 WITHOUT the `[B, T, Z]` logits: solve -> decoder -> the complete
        objective (`losses.mode_sep_total_loss`).  `union` = the `UnionBatch`, `y_union [B, T]` = zone id at snaps (-1 elsewhere).
        -> (total, parts)."""
        from .losses import mode_sep_total_loss
        E = self.config.emb_dim
        yb = self.integrate(self.initial_state(home_idx, work_idx, person_traits_raw), times_union).permute(1, 0, 2)
        return mode_sep_total_loss(self.config, self.decoder(yb[:, :, :E]), yb[:, :, E:2 * E], self.class_table, union, y_union, dist_mat)

    def forward(self, times_union, home_idx, work_idx, person_traits_raw) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        y0 = self.initial_state(home_idx, work_idx, person_traits_raw)
        return self.head(self.integrate(y0, times_union))
