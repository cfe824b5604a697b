"""CPU checks of the C ABI: libananke_b200.so builds for sm_100a (nvcc cross-compiles without a GPU), loads, and exports
exactly the symbols include/ananke_b200.h declares; the ctypes binding lists the same set.  No compute call is made."""
import ctypes as C
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from ananke_abm_b200 import build as b
    b.build(force=False)                      # in-tree, a no-op when up to date
    import ananke_abm_b200 as ab
    return ab.lib()


def _declared():
    text = (REPO / "include" / "ananke_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ab200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from ananke_abm_b200 import _lib
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported by the library"
    assert sorted(_lib.SIGNATURES) == names, (set(names) ^ set(_lib.SIGNATURES))


def test_host_only_queries(lib):
    from ananke_abm_b200 import _lib
    assert lib.ab200_abi_version() == 1
    assert lib.ab200_status_string(0) == b"ok" and b"dt" in lib.ab200_status_string(-6)
    ms = _lib.DriftDesc(64, 32, 128, 2, 0, 0, 0, 0, 0.0, 24.0)          # mode_sep drift
    lo = _lib.DriftDesc(16, 32, 128, 2, 1, 1, 12, 8, 1.0, 24.0)         # latent_ode drift
    assert lib.ab200_drift_param_count(C.byref(ms)) == 95168
    assert lib.ab200_drift_param_count(C.byref(lo)) == 76688
    assert lib.ab200_rk4_workspace_bytes(C.byref(ms), 1000, 97, 0) >= 95168 * 4
    assert lib.ab200_rk4_workspace_bytes(C.byref(ms), 1000, 97, 1) > 184 * 1024
    # the tensor-core stage path is instantiated for the mode_sep shape only
    assert lib.ab200_stage_image_bytes(C.byref(ms)) >= 2 * 210944
    assert lib.ab200_stage_image_bytes(C.byref(lo)) == 0
    assert lib.ab200_stage_spill_bytes(C.byref(ms), 10) == 10 * 389120   # 3,040 B per agent-stage x 128 agents
    assert lib.ab200_wgrad_partial_bytes(C.byref(ms)) > 0
    # argument validation happens before anything touches the device
    assert lib.ab200_rk4_forward(C.byref(ms), None, None, None, None, 10, 5, None, None, 0, 0, None) == -1
    assert lib.ab200_stage_forward(C.byref(lo), 1, 1, None, C.byref(C.c_int(0)), 10, None, None, None, 0, None) in (-1, -2)


def test_no_cpu_fallback():
    """the product path raises on CPU tensors instead of falling back"""
    import torch
    import ananke_abm_b200 as ab
    with pytest.raises(ab.Ab200Error):
        ab.odeint(lambda t, y: -y, torch.ones(3, 2), torch.linspace(0, 1, 3), method="rk4")
    # and nothing in the package imports the oracle
    for f in (REPO / "ananke_abm_b200").glob("*.py"):
        assert "oracle" not in f.read_text().replace("the oracle", "").replace("CPU oracle", ""), f
