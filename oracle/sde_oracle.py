"""CPU restatement of the SDE branch (TEST INFRASTRUCTURE ONLY -- nothing under ananke_abm_b200/ may import this).

Follows the reference's call  sdeint(sde, y0, ts, method="euler", dt=0.01)
(latent_ode/architecture/model.py:192-194, mode_sep/architecture/model.py:158-182; diagonal Ito noise,
`f` / `g` from the model: latent_ode/architecture/model.py:119-130, mode_sep/architecture/model.py:75-89) with the
fixed-step Euler-Maruyama scheme of torchsde 0.2.6 (`uv.lock:2928-2929`, not in the tree) [recall, unverifiable here]:
a grid t0, t0+dt, ... clipped at ts[-1]; every requested time is read off by LINEAR interpolation between the two grid
points around it (torchsde/_core/base_solver.py `integrate`).  PARITY UNPINNED for that stepping rule.

The Brownian increments cannot match torchsde's Brownian interval, so the noise follows the package's own specification
(csrc/sde_em.cu): Philox4x32-10, counter = (element group, step), key = seed, Box-Muller.  `philox4x32_10` is pinned on the
Random123 known-answer vectors (tests/test_oracle_sde.py).
"""
from __future__ import annotations

import numpy as np
import torch

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds; all arguments uint32 arrays (broadcastable)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0, k1 = np.asarray(k0, dtype=np.uint32), np.asarray(k1, dtype=np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = (k0 + _W0).astype(np.uint32)
            k1 = (k1 + _W1).astype(np.uint32)
    return c0, c1, c2, c3


def _box_muller(x0, x1):
    u1 = (x0.astype(np.float32) + np.float32(0.5)) * np.float32(2.3283064365386963e-10)
    u2 = (x1.astype(np.float32) + np.float32(0.5)) * np.float32(2.3283064365386963e-10)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    a = (np.float32(6.283185307179586) * u2).astype(np.float32)
    return (r * np.cos(a)).astype(np.float32), (r * np.sin(a)).astype(np.float32)


def normals(B: int, D: int, seed: int, step: int) -> np.ndarray:
    """xi [B, D] of one step (D % 4 == 0)."""
    q = np.arange(B * D // 4, dtype=np.uint64)
    x = philox4x32_10((q & _MASK).astype(np.uint32), (q >> np.uint64(32)).astype(np.uint32),
                      np.uint32(step & 0xFFFFFFFF), np.uint32(step >> 32), np.uint32(seed & 0xFFFFFFFF), np.uint32(seed >> 32))
    z0, z1 = _box_muller(x[0], x[1])
    z2, z3 = _box_muller(x[2], x[3])
    return np.stack([z0, z1, z2, z3], axis=-1).reshape(B, D)


def sdeint_euler(sde, y0: torch.Tensor, ts: torch.Tensor, dt: float, seed: int, with_grad: bool = False) -> torch.Tensor:
    """[len(ts), B, D]; `sde.f(t, y)`, `sde.g(t, y)` (diagonal noise) are the reference's own modules on the CPU.
    `with_grad`: record the loop with autograd (the reference trains through torchsde's solver ops)."""
    B, D = y0.shape
    t_list = [float(v) for v in ts.tolist()]
    out = [y0.clone()]
    curr_t, curr_y = t_list[0], y0.clone()
    prev_t, prev_y = curr_t, curr_y
    step = 0
    with torch.set_grad_enabled(bool(with_grad)):
        for out_t in t_list[1:]:
            while curr_t < out_t:
                next_t = min(curr_t + dt, t_list[-1])
                h = np.float32(next_t - curr_t)
                tt = torch.tensor(curr_t, dtype=torch.float32)
                f, g = sde.f(tt, curr_y), sde.g(tt, curr_y)
                xi = torch.from_numpy(normals(B, D, seed, step))
                prev_t, prev_y = curr_t, curr_y
                curr_y = curr_y + f * float(h) + g * float(np.sqrt(h)) * xi
                curr_t = next_t
                step += 1
            if curr_t == prev_t:
                out.append(curr_y.clone())
            else:
                w = (out_t - prev_t) / (curr_t - prev_t)
                out.append(prev_y + (curr_y - prev_y) * float(np.float32(w)))
    return torch.stack(out)
