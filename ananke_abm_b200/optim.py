"""Fused optimiser step: `clip_grad_norm_` + `torch.optim.Adam.step()` of the reference's training
loop (mode_sep/train/train.py:68,163-164) as two launches over flat fp32 buffers, without a host synchronisation.

    opt = FusedAdam(model.parameters(), lr=cfg.lr, weight_decay=cfg.weight_decay, max_grad_norm=cfg.grad_clip)
    loss.backward()
    flat = dist.allreduce_gradients(params)      # optional (N > 1): the flat buffer it returns can be passed to step()
    opt.step(flat)                               # or opt.step(): gathers p.grad itself

The parameters become views of one flat buffer (their values are preserved), so checkpoints (`state_dict`) are unchanged.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import _lib
from .dist import flatten_grads


class FusedAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: Optional[float] = None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdam got an empty parameter list")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise _lib.Ab200Error("FusedAdam: parameters must be CUDA tensors (no CPU path)")
        if any(p.dtype != torch.float32 or p.device != dev for p in self.params):
            raise ValueError("FusedAdam: all parameters must be float32 on one device")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.max_grad_norm = None if max_grad_norm is None else float(max_grad_norm)
        with torch.no_grad():
            self.flat = torch.cat([p.detach().reshape(-1) for p in self.params]).contiguous()
            off = 0
            for p in self.params:                      # parameters become views of the flat buffer
                n = p.numel()
                p.data = self.flat[off:off + n].view_as(p)
                off += n
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.step_count = 0

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self, flat_grad: Optional[torch.Tensor] = None) -> torch.Tensor:
        """-> total gradient norm before clipping (0-dim device tensor, like `clip_grad_norm_`)."""
        L = _lib.lib()
        g = flatten_grads(self.params) if flat_grad is None else flat_grad
        if g.numel() != self.flat.numel() or g.dtype != torch.float32 or not g.is_contiguous():
            raise ValueError("flat gradient buffer does not match the parameters")
        self.step_count += 1
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(L.ab200_grad_sumsq(g.data_ptr(), g.numel(), self.sumsq.data_ptr(), stream), "ab200_grad_sumsq")
        rc = L.ab200_adam_step(self.flat.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), g.numel(),
                               self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count,
                               self.max_grad_norm if self.max_grad_norm is not None else 0.0, self.sumsq.data_ptr(), stream)
        _lib.check(rc, "ab200_adam_step")
        return self.sumsq[0].sqrt().float()
