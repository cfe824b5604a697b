"""Build libananke_b200.so in-tree with nvcc for sm_100a (no torch headers involved: the library is a plain
C-ABI shared object loaded through ctypes).  `python -m ananke_abm_b200.build [--force]`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libananke_b200.so"
SOURCES = ["capi.cu", "rk4_f32.cu", "rk4_bwd_f32.cu", "rk_combine.cu", "umma_probe.cu", "rk4_tc.cu", "gat.cu",
           "stage_fwd_tc.cu", "stage_fwd2_tc.cu", "stage_bwd_tc.cu", "wgrad_tc.cu", "stage_elem.cu", "head_tc.cu", "head_bwd_tc.cu", "optim.cu", "sde_em.cu", "emb_losses.cu"]
OPTIONAL = []
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]
if os.environ.get("AB200_STAGE_TRACE") == "1":      # debug build: cycle trace of the stage kernels (scripts/trace_stage.py)
    NVCC_FLAGS.append("-DAB200_STAGE_TRACE")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    out = [CSRC / s for s in SOURCES]
    out += [CSRC / s for s in OPTIONAL if (CSRC / s).exists()]
    return out


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "ananke_b200.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in sources():
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src.name}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src.name}")
        objs.append(str(obj))
    (objdir / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
