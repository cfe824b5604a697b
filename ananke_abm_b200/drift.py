"""Structural recognition of the reference's drift modules, and the framework's own drift module.

`describe_drift(func)` looks at an `nn.Module` the way the solver seam receives it and, if it is the
"second-order residual-MLP drift" of either reference model, returns a `DriftSpec` (the C-ABI descriptor
plus the parameter tensors in `ab200_drift_desc` order).  No reference class is imported: recognition is by
attribute structure, so the UNMODIFIED reference modules are accepted as they are:

  mode_sep  `WrappedSDE(func=ODEFunc(...), emb_dim, context_dim)`
            /root/reference/src/ananke_abm/models/mode_sep/architecture/model.py:30-73
  latent    `ODEFunc(config, state_dim, position_dim, hidden_dim, num_residual_blocks)` (second-order branch)
            /root/reference/src/ananke_abm/models/latent_ode/architecture/model.py:19-117
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
from torch import nn

from . import _lib

# shapes instantiated in the CUDA library: (P, H, hid, n_res, res_act, potential)
_SUPPORTED = {(64, 32, 128, 2, 0, 0), (16, 32, 128, 2, 1, 1), (16, 32, 128, 2, 1, 0)}


@dataclass
class DriftSpec:
    desc: _lib.DriftDesc
    params: List[torch.Tensor]
    backward_precisions: Tuple[int, ...] = (_lib.PREC_F32,)

    @property
    def state_dim(self) -> int:
        return 2 * self.desc.pos_dim + self.desc.ctx_dim

    def flat_params(self) -> torch.Tensor:
        """Differentiable flatten in `ab200_drift_desc` order; autograd routes the flat gradient back."""
        return torch.cat([p.reshape(-1) for p in self.params])

    def tc_stage_supported(self) -> bool:
        """shape instantiated in the tensor-core stage kernels (stage_tc.cuh)"""
        d = self.desc
        return (d.pos_dim, d.ctx_dim, d.hidden, d.n_res, d.res_act, d.potential) == (64, 32, 128, 2, 0, 0)

    def supported(self) -> bool:
        d = self.desc
        return (d.pos_dim, d.ctx_dim, d.hidden, d.n_res, d.res_act, d.potential) in _SUPPORTED


def _parse_net(net) -> Optional[Tuple[List[torch.Tensor], int, int, int, int, int]]:
    """-> (params, in_dim, hidden, n_res, res_act, out_dim) for Linear,ReLU,[ResidualBlock]*,Linear stacks."""
    if not isinstance(net, nn.Sequential):
        return None
    mods = list(net)
    if len(mods) < 3 or not isinstance(mods[0], nn.Linear) or not isinstance(mods[1], nn.ReLU) \
            or not isinstance(mods[-1], nn.Linear):
        return None
    first, last = mods[0], mods[-1]
    if first.bias is None or last.bias is None:
        return None
    params = [first.weight, first.bias]
    act_code = None
    for blk in mods[2:-1]:
        inner, outer = getattr(blk, "net", None), getattr(blk, "activation", None)
        if not isinstance(inner, nn.Sequential) or len(inner) != 3 or outer is None:
            return None
        l1, a, l2 = inner
        if not (isinstance(l1, nn.Linear) and isinstance(l2, nn.Linear)) or type(a) is not type(outer):
            return None
        code = 0 if isinstance(a, nn.ReLU) else 1 if isinstance(a, nn.Tanh) else None
        if code is None or (act_code is not None and code != act_code):
            return None
        act_code = code
        if l1.weight.shape != (first.out_features, first.out_features) or l2.weight.shape != l1.weight.shape:
            return None
        params += [l1.weight, l1.bias, l2.weight, l2.bias]
    n_res = len(mods) - 3
    params += [last.weight, last.bias]
    return params, first.in_features, first.out_features, n_res, (act_code or 0), last.out_features


def describe_drift(func) -> Optional[DriftSpec]:
    if isinstance(func, SecondOrderDrift):
        return func.spec()
    if not isinstance(func, nn.Module):
        return None
    spec = None
    # --- mode_sep WrappedSDE: .func.net, .emb_dim, .context_dim
    inner = getattr(func, "func", None)
    if inner is not None and hasattr(func, "emb_dim") and hasattr(func, "context_dim") and hasattr(inner, "net"):
        parsed = _parse_net(inner.net)
        if parsed is not None:
            params, d_in, hid, n_res, act, d_out = parsed
            P, H = int(func.emb_dim), int(func.context_dim)
            if d_in == 2 * P + H + 2 and d_out == P:
                spec = DriftSpec(_lib.DriftDesc(P, H, hid, n_res, act, 0, 0, 0, 0.0, 24.0), params)
    # --- latent_ode ODEFunc: .net, .config, .state_dim, .position_dim (second-order branch only)
    elif hasattr(func, "net") and hasattr(func, "config") and hasattr(func, "position_dim") and hasattr(func, "state_dim"):
        cfg = func.config
        if getattr(cfg, "use_second_order_sde", False):
            parsed = _parse_net(func.net)
            if parsed is not None:
                params, d_in, hid, n_res, act, d_out = parsed
                P, H = int(func.position_dim), int(cfg.hidden_dim)
                if int(func.state_dim) == 2 * P and d_in == 2 * P + H + 2 and d_out == P:
                    ia = int(cfg.zone_embed_dim + cfg.purpose_feature_dim + getattr(func, "IS_MOVING_DIM", 0))
                    ib = int(cfg.zone_embed_dim + getattr(func, "IS_STATIONARY_DIM", 0))
                    spec = DriftSpec(_lib.DriftDesc(P, H, hid, n_res, act, 1, ia, ib, float(cfg.correction_strength), 24.0),
                                     params)
    if spec is None or not spec.supported():
        return None
    if any(p.dtype != torch.float32 or not p.is_cuda for p in spec.params):
        return None
    return spec


class _ResidualBlock(nn.Module):
    def __init__(self, dim: int, act: str):
        super().__init__()
        a = nn.ReLU if act == "relu" else nn.Tanh
        self.net = nn.Sequential(nn.Linear(dim, dim), a(), nn.Linear(dim, dim))
        self.activation = a()

    def forward(self, x):
        raise RuntimeError("evaluated only through the fused CUDA kernels")


class SecondOrderDrift(nn.Module):
    """dy/dt = [v, net([p, v, h, sin, cos]) (+ potential correction), 0] with the reference's parameter tree
    under `.net`, evaluated by the fused kernels only.  `forward(t, y)` is one `ab200_drift_eval` call."""
    _ab200_kernel_only = True     # forward() is a kernel call (ab200_drift_eval, with ab200_drift_vjp as its autograd backward)

    def __init__(self, pos_dim: int, ctx_dim: int, hidden: int = 128, n_res: int = 2, res_act: str = "relu",
                 potential: Optional[Tuple[int, int, float]] = None, period: float = 24.0):
        super().__init__()
        self.pos_dim, self.ctx_dim, self.hidden, self.n_res = pos_dim, ctx_dim, hidden, n_res
        self.res_act, self.potential, self.period = res_act, potential, period
        layers = [nn.Linear(2 * pos_dim + ctx_dim + 2, hidden), nn.ReLU()]
        layers += [_ResidualBlock(hidden, res_act) for _ in range(n_res)]
        layers.append(nn.Linear(hidden, pos_dim))
        self.net = nn.Sequential(*layers)

    def spec(self) -> DriftSpec:
        params, *_ = _parse_net(self.net)
        pa, pb, ps = self.potential if self.potential is not None else (0, 0, 0.0)
        desc = _lib.DriftDesc(self.pos_dim, self.ctx_dim, self.hidden, self.n_res, 0 if self.res_act == "relu" else 1,
                              0 if self.potential is None else 1, int(pa), int(pb), float(ps), float(self.period))
        return DriftSpec(desc, params)

    def forward(self, t, y):
        from .odeint import drift_apply
        return drift_apply(self.spec(), t, y)
