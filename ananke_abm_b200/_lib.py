"""ctypes binding of libananke_b200.so (the C ABI declared in include/ananke_b200.h).

The product path has no CPU fallback: if the shared library is missing or fails to load, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libananke_b200.so"

PREC_F32, PREC_BF16, PREC_BF16X3 = 0, 1, 2
PRECISIONS = {"f32": PREC_F32, "fp32": PREC_F32, "float32": PREC_F32, "bf16": PREC_BF16, "bf16x3": PREC_BF16X3}


class DriftDesc(C.Structure):
    """Mirror of `ab200_drift_desc`."""
    _fields_ = [
        ("pos_dim", C.c_int32), ("ctx_dim", C.c_int32), ("hidden", C.c_int32), ("n_res", C.c_int32),
        ("res_act", C.c_int32), ("potential", C.c_int32), ("pot_idx_a", C.c_int32), ("pot_idx_b", C.c_int32),
        ("pot_strength", C.c_float), ("time_period", C.c_float),
    ]

    def key(self):
        return tuple(getattr(self, f) for f, _ in self._fields_)


class Ab200Error(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None

_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
_dp = C.POINTER(DriftDesc)

# name -> (restype, argtypes); must list every symbol include/ananke_b200.h declares (tests check this)
SIGNATURES = {
    "ab200_abi_version": (C.c_int, []),
    "ab200_status_string": (C.c_char_p, [C.c_int]),
    "ab200_last_cuda_error": (C.c_char_p, []),
    "ab200_drift_param_count": (_i64, [_dp]),
    "ab200_rk4_workspace_bytes": (_sz, [_dp, _i64, _i32, _i32]),
    "ab200_rk4_forward": (C.c_int, [_dp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _sz, _i32, _vp]),
    "ab200_rk4_backward_workspace_bytes": (_sz, [_dp, _i64, _i32, _i32]),
    "ab200_rk4_backward": (C.c_int, [_dp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _sz, _i32, _vp]),
    "ab200_drift_eval_workspace_bytes": (_sz, [_dp, _i64, _i32]),
    "ab200_drift_eval": (C.c_int, [_dp, _vp, _f32, _vp, _i64, _vp, _vp, _sz, _i32, _vp]),
    "ab200_drift_vjp_workspace_bytes": (_sz, [_dp, _i64]),
    "ab200_drift_vjp": (C.c_int, [_dp, _vp, _f32, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "ab200_emb_losses_forward": (C.c_int, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32,
                                           _f32, _f32, _f32, _f32, _vp, _vp]),
    "ab200_emb_losses_backward": (C.c_int, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32,
                                            _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _vp]),
    "ab200_rk_stage_combine": (C.c_int, [_vp, _vp, _vp, _i32, _f32, _vp, _i64, _vp]),
    "ab200_rk_combine_errnorm": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _f32, _f32, _f32, _vp, _vp, _i64, _vp]),
    "ab200_head_workspace_bytes": (_sz, [_i32, _i32]),
    "ab200_sde_euler_step": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _i32, _f32, C.c_uint64, C.c_uint64, _vp, _vp, _vp]),
    "ab200_grad_sumsq": (C.c_int, [_vp, _i64, _vp, _vp]),
    "ab200_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _f32, _vp, _vp]),
    "ab200_head_argmax": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _f32, _vp, _vp, _vp, _sz, _vp]),
    "ab200_head_ce_forward": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ab200_head_ce_backward_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "ab200_head_ce_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _f32, _vp, _vp, _vp, _sz, _vp]),
    "ab200_head_ce_backward_status": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "ab200_gat_forward": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp,
                                    _vp, _vp]),
    "ab200_gat_backward_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "ab200_gat_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp, _vp,
                                     _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ab200_rows_block": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "ab200_rows_unblock": (C.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "ab200_stage_image_bytes": (_sz, [_dp]),
    "ab200_stage_pack": (C.c_int, [_dp, _vp, _vp, _sz, _vp]),
    "ab200_stage_forward": (C.c_int, [_dp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _vp]),
    "ab200_stage_forward_fused": (C.c_int, [_dp, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp, _vp, _i32, _vp]),
    "ab200_stage_forward_fused_save": (C.c_int, [_dp, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp, _vp, _vp, _i32, _vp]),
    "ab200_dopri5_attempt": (C.c_int, [_dp, _vp, _vp, _vp, C.c_double, C.c_double, _i64, _vp, _vp, _f32, _f32, _i32, _vp, _i32, _vp]),
    "ab200_stage_xblob_bytes": (_sz, [_dp, _i64, _i32]),
    "ab200_dopri5_dense_rows": (C.c_int, [_dp, _vp, _vp, C.c_double, _i32, _vp, _i64, _vp, _vp]),
    "ab200_stage_spill_bytes": (_sz, [_dp, _i32]),
    "ab200_wgrad_partial_bytes": (_sz, [_dp]),
    "ab200_stage_backward": (C.c_int, [_dp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _sz, _i32, _i32, _vp, _vp]),
    "ab200_stage_backward_fused": (C.c_int, [_dp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _sz, _i32, _i32,
                                             _vp, _vp, _i32, _vp, _vp, _vp]),
    "ab200_adjoint_gather": (C.c_int, [_dp, _vp, _vp, _i32, _vp, _i64, _vp, _vp]),
    "ab200_adjoint_gather_upstream": (C.c_int, [_dp, _vp, _vp, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ab200_pv_combine_backward_multi": (C.c_int, [_dp, _vp, _i32, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _i32, _i32, _vp, _i32, _vp]),
    "ab200_stage_upstream": (C.c_int, [_dp, _vp, _vp, _i32, _vp, _vp, _i64, _vp, _vp]),
    "ab200_wgrad_accumulate": (C.c_int, [_dp, _vp, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _vp]),
    "ab200_wgrad_finalize": (C.c_int, [_dp, _vp, _vp, _vp]),
    "ab200_stage_status_offset": (C.c_int, [_dp, _vp, _vp]),
    "ab200_pv_combine": (C.c_int, [_dp, _vp, _vp, _i32, _f32, _vp, _vp, _i64, _vp, _vp]),
    "ab200_pv_combine_rowmajor": (C.c_int, [_dp, _vp, _vp, _i32, _f32, _vp, _vp, _i64, _vp, _vp]),
    "ab200_pv_combine_rowmajor_multi": (C.c_int, [_dp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _vp]),
    "ab200_pv_combine_backward": (C.c_int, [_dp, _vp, _i32, _f32, _vp, _vp, _i64, _vp, _vp, _i32, _vp]),
    "ab200_debug_umma_probe": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
}


def lib() -> C.CDLL:
    """Load (once) and return the shared library.  Raises if it is not built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if os.environ.get("ANANKE_B200_AUTOBUILD", "0") == "1":
            from . import build as _b
            _b.build()
        else:
            raise Ab200Error(
                f"{LIB_PATH} is missing: build it with `python -m ananke_abm_b200.build` "
                "(the CUDA path is the only path; there is no CPU fallback)")
    handle = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)          # AttributeError here == header/library drift
        fn.restype = res
        fn.argtypes = args
    if handle.ab200_abi_version() != 1:
        raise Ab200Error("ABI version mismatch between the Python binding and libananke_b200.so")
    _lib = handle
    return handle


LAUNCHES = 0     # C-ABI calls that launched kernels since the counter was last reset (bench.py reports it)


def check(status: int, what: str) -> None:
    global LAUNCHES
    LAUNCHES += 1
    if status == 0:
        return
    L = lib()
    msg = L.ab200_status_string(status).decode()
    if status == -4:
        msg += ": " + L.ab200_last_cuda_error().decode()
    if status == -5:
        # same wording torchdiffeq uses (odeint.py `_check_inputs`)
        raise AssertionError("t must be strictly increasing or decreasing")
    if status == -6:
        raise AssertionError("underflow in dt")
    if status == -7:
        raise AssertionError("max_num_steps exceeded")
    raise Ab200Error(f"{what}: {msg} (status {status})")
