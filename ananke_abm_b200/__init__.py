"""ananke_abm_b200 -- B200-native GAT-ODE hot path of bobkatla/ananke_abm behind the reference's interfaces.

Public surface:
    odeint, odeint_adjoint      torchdiffeq-compatible solver seam (fused sm_100a kernels underneath)
    ModeSepModel, ModeSepConfig reference module signature (mode_sep)
    SecondOrderDrift            the drift module family the kernels implement
The CUDA shared library (libananke_b200.so, C ABI in include/ananke_b200.h) is the only compute path.
"""
from ._lib import Ab200Error, lib  # noqa: F401
from .drift import SecondOrderDrift, describe_drift  # noqa: F401
from .odeint import odeint, odeint_adjoint, set_default_precision, drift_eval, drift_vjp  # noqa: F401
from .mode_sep import ModeSepConfig, ModeSepModel  # noqa: F401
from .gnn_embed import GATEmbed, gnn_embed  # noqa: F401
from .graph import ZoneCSR, build_zone_csr  # noqa: F401
from .run import GATODEModel  # noqa: F401
from .latent_ode import GenerativeODE, GenerativeODEConfig  # noqa: F401
from .batching import UnionBatch, build_union_batch, unify_and_interpolate_batch  # noqa: F401
from .losses import (ce_and_expected_distance_at_snaps_fused, ce_at_snaps_fused, head_ce_rows, head_loss_rows,  # noqa: F401
                     emb_loss_terms, mode_sep_total_loss)
from .optim import FusedAdam  # noqa: F401
from .sdeint import sdeint  # noqa: F401
from . import inference, run, batching, latent_ode, stage  # noqa: F401

__version__ = "0.1.0"
