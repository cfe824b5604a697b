"""ORACLE (test infrastructure, not product code) -- plain-PyTorch fp32 restatement of the two reference
models whose ODE right-hand side is on the hot path.  Runs on CPU; nothing here is ever timed as the
product and nothing under `ananke_abm_b200/` imports it.

Pinned against the reference itself: tests/golden/make_golden.py imports the UNMODIFIED reference modules
from /root/reference/src (with `oracle.torchdiffeq_oracle` standing in for the absent torchdiffeq) and
freezes their outputs; tests/test_oracle_models.py checks this restatement against those vectors, and --
when /root/reference is present -- against the live reference modules, parameter for parameter.

Reference anchors (all under /root/reference/src/ananke_abm/models/):
  mode_sep/architecture/model.py:16-27   ResidualBlock  (x -> relu(x + L2(relu(L1 x))))
  mode_sep/architecture/model.py:30-38   ODEFunc.net    (Linear(2E+H+2, hid), ReLU, nres x ResidualBlock, Linear(hid, E))
  mode_sep/architecture/model.py:56-73   WrappedSDE.forward  (dp=v, dv=net([p,v,h,sin,cos]), dh=0)
  mode_sep/architecture/model.py:92-201  ModeSepModel   (tables, context encoder, decoder, forward)
  latent_ode/architecture/model.py:9-17  ResidualBlock  (tanh flavour)
  latent_ode/architecture/model.py:56-117 ODEFunc       (MLP + potential-gradient correction)
  latent_ode/architecture/model.py:132-220 GenerativeODE
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from . import torchdiffeq_oracle as tdq


# --------------------------------------------------------------------------------------------------
# configuration mirrors (only the fields the hot path reads)
# --------------------------------------------------------------------------------------------------
@dataclass
class ModeSepDims:           # mode_sep/config.py:16-28,41
    emb_dim: int = 64
    context_dim: int = 32
    zone_emb_dim: int = 8
    hidden_dim: int = 128
    num_res_blocks: int = 2
    ode_method: str = "rk4"
    rtol: float = 1e-5
    atol: float = 1e-5
    softmax_tau: float = 0.2


@dataclass
class LatentDims:            # latent_ode/config.py:21-35,57 ; feature dims from data_generator/feature_engineering.py
    hidden_dim: int = 32
    encoder_hidden_dim: int = 64
    ode_hidden_dim: int = 128
    zone_embed_dim: int = 8
    purpose_feature_dim: int = 4
    mode_feature_dim: int = 4
    num_residual_blocks: int = 2
    correction_strength: float = 1.0
    use_second_order_sde: bool = True       # latent_ode/config.py: the only branch restated (and the reference default)
    ode_method: str = "dopri5"
    num_modes: int = 5
    num_purposes: int = 6


# --------------------------------------------------------------------------------------------------
# module skeletons with the reference's parameter names, so state_dicts are interchangeable and the
# default initialisation under a given torch seed consumes the RNG stream in the same order
# --------------------------------------------------------------------------------------------------
class _Res(nn.Module):
    def __init__(self, dim: int, act: str):
        super().__init__()
        a = nn.ReLU if act == "relu" else nn.Tanh
        self.net = nn.Sequential(nn.Linear(dim, dim), a(), nn.Linear(dim, dim))
        self.activation = a()

    def forward(self, x):
        return self.activation(x + self.net(x))


def _drift_net(d_in: int, hid: int, nres: int, d_out: int, res_act: str) -> nn.Sequential:
    layers = [nn.Linear(d_in, hid), nn.ReLU()]
    layers += [_Res(hid, res_act) for _ in range(nres)]
    layers.append(nn.Linear(hid, d_out))
    return nn.Sequential(*layers)


class _Holder(nn.Module):
    pass


class OracleWrappedSDE(nn.Module):
    """The reference's `WrappedSDE(func=ODEFunc(...), emb_dim, context_dim)` (mode_sep/architecture/model.py:49-73): same
    attribute structure (`.func.net`, `.emb_dim`, `.context_dim`) and the same eager forward, so that whatever consumes the
    solver seam sees what it would see from the unmodified reference class.  `calls` counts eager evaluations (tests use it to
    prove that a drop-in solver evaluated the drift on its own kernels instead)."""

    def __init__(self, net: nn.Sequential, emb_dim: int, context_dim: int):
        super().__init__()
        self.func = _Holder()
        self.func.net = net
        self.emb_dim, self.context_dim = emb_dim, context_dim
        self.calls = 0

    def forward(self, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.calls += 1
        E, H = self.emb_dim, self.context_dim
        p, v, h = torch.split(y, [E, E, H], dim=-1)
        s = torch.sin(t * 2 * torch.pi / 24.0).expand(y.shape[0])
        c = torch.cos(t * 2 * torch.pi / 24.0).expand(y.shape[0])
        a = self.func.net(torch.cat([p, v, h, torch.stack([s, c], dim=-1)], dim=-1))
        return torch.cat([v, a, torch.zeros_like(h)], dim=-1)


class OracleLatentFunc(nn.Module):
    """The reference's latent `ODEFunc(config, state_dim, position_dim, ...)`, second-order branch
    (latent_ode/architecture/model.py:19-117): `.net`, `.config`, `.state_dim`, `.position_dim` and the eager forward with the
    potential-gradient correction in closed form (see OracleLatentODE.rhs)."""
    IS_MOVING_DIM = 0
    IS_STATIONARY_DIM = 0

    def __init__(self, dims, net: nn.Sequential, position_dim: int):
        super().__init__()
        self.config = dims
        self.net = net
        self.position_dim, self.state_dim = position_dim, 2 * position_dim
        self.calls = 0

    def forward(self, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.calls += 1
        d, P = self.config, self.position_dim
        p, v, h = torch.split(y, [P, P, d.hidden_dim], dim=-1)
        s = torch.sin(t * 2 * torch.pi / 24).expand(y.shape[0])
        c = torch.cos(t * 2 * torch.pi / 24).expand(y.shape[0])
        dv = self.net(torch.cat([p, v, h, torch.stack([s, c], dim=-1)], dim=-1))
        i_purp = d.zone_embed_dim
        i_mode = d.zone_embed_dim + d.purpose_feature_dim
        a = torch.sigmoid(p[:, i_mode])
        b = torch.sigmoid(p[:, i_purp])
        r = 2.0 * (a + b - 1.0)
        corr = torch.zeros_like(dv)
        corr[:, i_mode] = -r * a * (1 - a)
        corr[:, i_purp] = -r * b * (1 - b)
        dv = dv + d.correction_strength * corr
        return torch.cat([v, dv, torch.zeros_like(h)], dim=-1)


class OracleModeSep(nn.Module):
    """Same parameter tree as the reference ModeSepModel (mode_sep/architecture/model.py:92-136)."""

    def __init__(self, Z: int, dims: Optional[ModeSepDims] = None):
        super().__init__()
        d = dims or ModeSepDims()
        self.dims, self.Z = d, Z
        E, H = d.emb_dim, d.context_dim
        self.class_table = nn.Parameter(torch.empty(Z, E))
        nn.init.xavier_uniform_(self.class_table)
        self.zone_embed = nn.Embedding(Z, d.zone_emb_dim)
        self.context_encoder = nn.Sequential(nn.Linear(2 + 2 * d.zone_emb_dim, d.hidden_dim), nn.ReLU(),
                                             nn.Linear(d.hidden_dim, H))
        self.odefunc = OracleWrappedSDE(_drift_net(2 * E + H + 2, d.hidden_dim, d.num_res_blocks, E, "relu"), E, H)
        self.decoder = nn.Sequential(nn.Linear(E, d.hidden_dim), nn.ReLU(), nn.Linear(d.hidden_dim, E))

    # ---- WrappedSDE.forward (mode_sep/architecture/model.py:56-73)
    def rhs(self, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return self.odefunc(t, y)

    # ---- ModeSepModel.forward initial state (model.py:150-155)
    def initial_state(self, home_idx, work_idx, traits) -> torch.Tensor:
        p0 = self.class_table.detach()[home_idx]
        raw = torch.cat([traits, self.zone_embed(home_idx), self.zone_embed(work_idx)], dim=-1)
        h = self.context_encoder(raw)
        return torch.cat([p0, torch.zeros_like(p0), h], dim=-1)

    # ---- head (model.py:192-199)
    def head(self, y_path: torch.Tensor):
        E, H = self.dims.emb_dim, self.dims.context_dim
        yb = y_path.permute(1, 0, 2)
        p_t, v_t, _ = torch.split(yb, [E, E, H], dim=-1)
        pred = self.decoder(p_t)
        tab = self.class_table / (self.class_table.norm(dim=-1, keepdim=True) + 1e-8)
        en = pred / (pred.norm(dim=-1, keepdim=True) + 1e-8)
        logits = torch.einsum("bte,ze->btz", en, tab) / self.dims.softmax_tau
        return pred, logits, v_t

    def solve(self, y0, times):
        # the reference's call shape (mode_sep/architecture/model.py:184-191): the drift MODULE goes through the solver seam
        return tdq.odeint(self.odefunc, y0, times, method=self.dims.ode_method, rtol=self.dims.rtol,
                          atol=self.dims.atol)

    def forward(self, times_union, home_idx, work_idx, person_traits_raw):
        y0 = self.initial_state(home_idx, work_idx, person_traits_raw)
        return self.head(self.solve(y0, times_union))


class _OracleGAT(nn.Module):
    """PyG GATConv parameter tree (`lin.weight`, `att_src`, `att_dst`, `bias`) evaluated by oracle.gat_oracle.gat_edges."""

    def __init__(self, in_channels: int, out_channels: int, heads: int):
        super().__init__()
        from . import gat_oracle as go
        self.heads, self.out_channels = heads, out_channels
        self.lin = _Holder()
        self.lin.weight = nn.Parameter(go.glorot_(torch.empty(heads * out_channels, in_channels)))
        self.att_src = nn.Parameter(go.glorot_(torch.empty(1, heads, out_channels)))
        self.att_dst = nn.Parameter(go.glorot_(torch.empty(1, heads, out_channels)))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels))

    def forward(self, x, edges):
        from . import gat_oracle as go
        return go.gat_edges(x, edges, self.lin.weight, self.att_src.view(-1), self.att_dst.view(-1), self.bias, self.heads,
                            self.out_channels, concat=True)


class OracleGATODE(OracleModeSep):
    """The GAT-ODE of north_star in plain PyTorch (SURVEY.md Open Question 1, Reading A): ModeSepModel with its two learnable
    zone lookups (`class_table[Z, E]`, `zone_embed[Z, 8]`, mode_sep/architecture/model.py:101-102) produced by graph
    attention layers over the zone features (PyG GATConv semantics, gat_oracle.py).  Same parameter names as
    `ananke_abm_b200.run.GATODEModel`, so state dicts are interchangeable.  Runs on whatever device its tensors are on:
    the CPU arm and the eager-PyTorch-on-GPU arm of bench.py time this module."""

    def __init__(self, num_zone_features: int, heads: int = 4, dims: Optional[ModeSepDims] = None):
        super().__init__(1, dims)
        del self.class_table, self.zone_embed
        d = self.dims
        self.table_gat = _OracleGAT(num_zone_features, d.emb_dim // heads, heads)
        self.zone_gat = _OracleGAT(num_zone_features, d.zone_emb_dim, 1)

    def zone_tables(self, zone_features, edges):
        return self.table_gat(zone_features, edges), self.zone_gat(zone_features, edges)

    def initial_state(self, class_table, zone_embed, home_idx, work_idx, traits) -> torch.Tensor:      # noqa: signature differs on purpose
        p0 = class_table.detach()[home_idx]
        raw = torch.cat([traits, zone_embed[home_idx], zone_embed[work_idx]], dim=-1)
        return torch.cat([p0, torch.zeros_like(p0), self.context_encoder(raw)], dim=-1)


class OracleLatentODE(nn.Module):
    """Same parameter tree as the reference GenerativeODE (latent_ode/architecture/model.py:132-165),
    ODE branch only (`enable_sde=False`); the h0 reparameterisation noise is an explicit argument."""

    def __init__(self, person_feat_dim: int, num_zone_features: int, dims: Optional[LatentDims] = None):
        super().__init__()
        d = dims or LatentDims()
        self.dims = d
        self.zone_feature_encoder = nn.Linear(num_zone_features, d.zone_embed_dim)
        enc_in = person_feat_dim + 2 * d.zone_embed_dim + d.purpose_feature_dim + d.mode_feature_dim
        self.encoder = nn.Sequential(nn.Linear(enc_in, d.encoder_hidden_dim), nn.ReLU(),
                                     nn.Linear(d.encoder_hidden_dim, 2 * d.hidden_dim))
        self.position_dim = d.zone_embed_dim + d.purpose_feature_dim + d.mode_feature_dim
        self.state_dim = 2 * self.position_dim
        self.ode_func = OracleLatentFunc(d, _drift_net(self.state_dim + d.hidden_dim + 2, d.ode_hidden_dim,
                                                       d.num_residual_blocks, self.position_dim, "tanh"), self.position_dim)
        self.decoder_loc = nn.Linear(d.zone_embed_dim, d.zone_embed_dim)
        self.decoder_purpose = nn.Linear(d.purpose_feature_dim, d.num_purposes)
        self.decoder_mode = nn.Linear(d.mode_feature_dim, d.num_modes)

    # ---- ODEFunc.forward, second-order branch (latent_ode/architecture/model.py:77-117).  The
    # autograd.grad of the potential (model.py:56-74,93-95) is written out in closed form:
    #   U = sum_b (a + b - 1)^2,  a = sigmoid(p[mode dim 0]),  b = sigmoid(p[purpose dim 0])
    #   -dU/dp[mode0] = -2 (a+b-1) a (1-a) ;  -dU/dp[purp0] = -2 (a+b-1) b (1-b)
    # applied only when any potential term is > 0 (the reference's `torch.any(potential > 0)` branch,
    # which is a no-op otherwise because the gradient is then exactly zero too).
    def rhs(self, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return self.ode_func(t, y)

    def initial_state(self, person_features, home_zone_features, work_zone_features, purp0, mode0, eps):
        d = self.dims
        home = self.zone_feature_encoder(home_zone_features)
        work = self.zone_feature_encoder(work_zone_features)
        mu, logvar = self.encoder(torch.cat([person_features, home, work, purp0, mode0], dim=-1)).split(
            d.hidden_dim, dim=-1)
        h0 = mu + torch.exp(0.5 * logvar) * eps
        p0 = torch.cat([home, purp0, mode0], dim=-1)
        return torch.cat([p0, torch.zeros_like(p0), h0], dim=-1), mu, logvar

    def solve(self, y0, times, rtol=1e-7, atol=1e-9):
        # the reference's call shape (latent_ode/architecture/model.py:192,196)
        return tdq.odeint(self.ode_func, y0, times, method=self.dims.ode_method, rtol=rtol, atol=atol,
                          options={"dtype": torch.float32})

    def head(self, y_path, all_zone_features):
        d, P = self.dims, self.position_dim
        yb = y_path.permute(1, 0, 2)
        pp = yb[..., :P]
        loc, purp, mode = torch.split(pp, [d.zone_embed_dim, d.purpose_feature_dim, d.mode_feature_dim], dim=-1)
        cand = self.zone_feature_encoder(all_zone_features)
        loc_logits = torch.einsum("bsd,zd->bsz", self.decoder_loc(loc), cand)
        return loc_logits, loc, self.decoder_purpose(purp), self.decoder_mode(mode), purp, mode

    def forward(self, person_features, home_zone_features, work_zone_features, purp0, mode0, times,
                all_zone_features, eps):
        y0, mu, logvar = self.initial_state(person_features, home_zone_features, work_zone_features, purp0,
                                            mode0, eps)
        out = self.head(self.solve(y0, times), all_zone_features)
        return (*out, mu, logvar)


# --------------------------------------------------------------------------------------------------
# the generic "second-order residual-MLP drift" the CUDA kernels implement, written out functionally
# --------------------------------------------------------------------------------------------------
def drift_weights(net: nn.Sequential) -> Dict[str, torch.Tensor]:
    """Flatten `ODEFunc.net` (mode_sep model.py:34-38 / latent model.py:46-51) into named tensors."""
    mods = list(net)
    out = {"w_in": mods[0].weight, "b_in": mods[0].bias, "w_out": mods[-1].weight, "b_out": mods[-1].bias}
    res = [m for m in mods if hasattr(m, "net") and hasattr(m, "activation")]
    for i, r in enumerate(res):
        out[f"w_r{i}a"], out[f"b_r{i}a"] = r.net[0].weight, r.net[0].bias
        out[f"w_r{i}b"], out[f"b_r{i}b"] = r.net[2].weight, r.net[2].bias
    return out


def drift_accel(x: torch.Tensor, w: Dict[str, torch.Tensor], nres: int, res_act: str) -> torch.Tensor:
    act = torch.relu if res_act == "relu" else torch.tanh
    z = torch.relu(x @ w["w_in"].T + w["b_in"])
    for i in range(nres):
        u = act(z @ w[f"w_r{i}a"].T + w[f"b_r{i}a"])
        z = act(z + u @ w[f"w_r{i}b"].T + w[f"b_r{i}b"])
    return z @ w["w_out"].T + w["b_out"]


def rk4_38_path(rhs, y0: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """Thin alias used by tests that want the fixed-grid path without the odeint argument plumbing."""
    return tdq.odeint(rhs, y0, t, method="rk4")
