"""`run` -- the GAT-ODE model assembled from the pieces: graph-attention zone embedding -> per-agent initial state
-> fused Runge-Kutta integration of the drift net -> decoder / cosine head.  (Namespace promised by the reference
README, /root/reference/README.md:57,78-80; wiring per SURVEY.md Open Question 1, Reading A: the GAT produces the
shared zone tables, agents evolve [p, v, h] by the drift net and read the tables at y0 and in the head.)

`GATODEModel` keeps `ModeSepModel`'s forward signature and outputs
(/root/reference/src/ananke_abm/models/mode_sep/architecture/model.py:138-201); the learnable `class_table[Z,E]` and
`zone_embed[Z,8]` lookups are replaced by GAT layers over the zone features (4 heads x 16 or 1 head x 64 for the
class table, 1 head x 8 for the zone embedding -- the shape hooks of SURVEY.md App. B).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from .gnn_embed import GATEmbed
from .graph import ZoneCSR
from .mode_sep import ModeSepConfig, ODEFunc, WrappedSDE, _solver_options
from .odeint import odeint, odeint_adjoint


class GATODEModel(nn.Module):
    def __init__(self, num_zone_features: int, config: ModeSepConfig, heads: int = 4):
        super().__init__()
        self.config, self.heads = config, heads
        E, H = config.emb_dim, config.context_dim
        assert E % heads == 0
        self.table_gat = GATEmbed(num_zone_features, E // heads, heads=heads, concat=True)
        self.zone_gat = GATEmbed(num_zone_features, config.zone_emb_dim, heads=1, concat=True)
        self.context_encoder = nn.Sequential(nn.Linear(2 + 2 * config.zone_emb_dim, config.hidden_dim), nn.ReLU(),
                                             nn.Linear(config.hidden_dim, H))
        self.odefunc = WrappedSDE(ODEFunc(E, H, config.hidden_dim, config.num_res_blocks), E, H)
        self.decoder = nn.Sequential(nn.Linear(E, config.hidden_dim), nn.ReLU(), nn.Linear(config.hidden_dim, E))

    def zone_tables(self, zone_features: torch.Tensor, graph: ZoneCSR) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.table_gat(zone_features, graph), self.zone_gat(zone_features, graph)

    def initial_state(self, class_table, zone_embed, home_idx, work_idx, traits) -> torch.Tensor:
        p0 = class_table.detach()[home_idx]                    # model.py:150-151: the table is detached for p0
        raw = torch.cat([traits, zone_embed[home_idx], zone_embed[work_idx]], dim=-1)
        h = self.context_encoder(raw)
        return torch.cat([p0, torch.zeros_like(p0), h], dim=-1)

    def integrate(self, y0, times_union) -> torch.Tensor:
        if getattr(self.config, "adjoint", False):                                       # the odeint_adjoint seam on request
            extra = {}
            if getattr(self.config, "adjoint_options", None) and getattr(self.config, "adjoint_mode", "continuous") != "discrete":
                extra["adjoint_options"] = dict(self.config.adjoint_options)             # e.g. {"norm": "seminorm"} (torchdiffeq adjoint.py)
            return odeint_adjoint(self.odefunc, y0, times_union, method=self.config.ode_method, rtol=self.config.rtol,
                                  atol=self.config.atol, options=_solver_options(self.config), **extra)
        return odeint(self.odefunc, y0, times_union, method=self.config.ode_method, rtol=self.config.rtol,
                      atol=self.config.atol, options=_solver_options(self.config))

    def head(self, y_path, class_table):
        E, H = self.config.emb_dim, self.config.context_dim
        yb = y_path.permute(1, 0, 2)
        p_t, v_t, _ = torch.split(yb, [E, E, H], dim=-1)
        pred_emb = self.decoder(p_t)
        table_norm = class_table / (class_table.norm(dim=-1, keepdim=True) + 1e-8)
        emb_norm = pred_emb / (pred_emb.norm(dim=-1, keepdim=True) + 1e-8)
        return pred_emb, torch.einsum("bte,ze->btz", emb_norm, table_norm) / self.config.softmax_tau, v_t

    def training_loss(self, times_union, home_idx, work_idx, person_traits_raw, zone_features, graph: ZoneCSR, union, y_union, dist_mat):
        """One training forward WITHOUT the `[B, T, Z]` logits: zone tables -> initial state -> solve -> decoder -> the complete
        mode_sep objective (`losses.mode_sep_total_loss`: fused tensor-core head for cross entropy / expected distance, fused
        pass for the embedding-space terms).  -> (total, parts); call `.backward()` on `total`
        (mode_sep/train/train.py:94-162 with the GAT tables in place of the learnable lookups)."""
        from .losses import mode_sep_total_loss
        E = self.config.emb_dim
        class_table, zone_embed = self.zone_tables(zone_features, graph)
        y0 = self.initial_state(class_table, zone_embed, home_idx, work_idx, person_traits_raw)
        yb = self.integrate(y0, times_union).permute(1, 0, 2)
        return mode_sep_total_loss(self.config, self.decoder(yb[:, :, :E]), yb[:, :, E:2 * E], class_table, union, y_union, dist_mat)

    def forward(self, times_union, home_idx, work_idx, person_traits_raw, zone_features, graph: ZoneCSR):
        class_table, zone_embed = self.zone_tables(zone_features, graph)
        y0 = self.initial_state(class_table, zone_embed, home_idx, work_idx, person_traits_raw)
        return self.head(self.integrate(y0, times_union), class_table)


def integrate(model, *args, **kwargs):
    """Functional alias: `run.integrate(model, times, home, work, traits[, zone_features, graph])`."""
    return model(*args, **kwargs)
