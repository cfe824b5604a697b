"""`gnn_embed` -- graph-attention embedding of the zone graph (the namespace the reference README promises,
/root/reference/README.md:57,78-80; no implementation exists in the reference).

`GATEmbed` is a drop-in for PyG `GATConv(in_channels, out_channels, heads, concat, negative_slope=0.2,
add_self_loops=True, bias=True)`: same parameter names (`lin.weight`, `att_src`, `att_dst`, `bias`), same glorot/zero
initialisation, same output layout.  Shape hooks: 1 head x 8 replaces the `zone_feature_encoder`
output [Z, 8] (latent_ode/architecture/model.py:141,171); 4 heads x 16 concatenated replaces `class_table` [Z, 64]
(mode_sep/architecture/model.py:101).  Forward and backward run in the fused CSR kernels of libananke_b200.so.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
from torch import nn

from . import _lib
from .graph import ZoneCSR, build_zone_csr


def _sp() -> int:
    return torch.cuda.current_stream().cuda_stream


class _GATFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, att_src, att_dst, bias, csr: ZoneCSR, heads: int, F_out: int, concat: bool, slope: float):
        L = _lib.lib()
        if not x.is_cuda:
            raise _lib.Ab200Error("GATEmbed: inputs must be CUDA tensors (no CPU path)")
        Z, F_in = x.shape
        HF = heads * F_out
        xc, Wc = x.contiguous().float(), W.contiguous().float()
        as_c, ad_c = att_src.contiguous().float().view(-1), att_dst.contiguous().float().view(-1)
        bc = None if bias is None else bias.contiguous().float()
        dev = x.device
        out = torch.empty((Z, HF if concat else F_out), dtype=torch.float32, device=dev)
        xw = torch.empty((Z, HF), dtype=torch.float32, device=dev)
        a_s = torch.empty((Z, heads), dtype=torch.float32, device=dev)
        a_d = torch.empty((Z, heads), dtype=torch.float32, device=dev)
        alpha = torch.empty((csr.nnz, heads), dtype=torch.float32, device=dev)
        rc = L.ab200_gat_forward(csr.rowptr.data_ptr(), csr.col.data_ptr(), Z, csr.nnz, xc.data_ptr(), F_in, Wc.data_ptr(),
                                 as_c.data_ptr(), ad_c.data_ptr(), None if bc is None else bc.data_ptr(), heads, F_out,
                                 1 if concat else 0, float(slope), out.data_ptr(), xw.data_ptr(), a_s.data_ptr(), a_d.data_ptr(),
                                 alpha.data_ptr(), _sp())
        _lib.check(rc, "ab200_gat_forward")
        ctx.save_for_backward(xc, Wc, as_c, ad_c, xw, a_s, a_d, alpha)
        ctx.csr, ctx.cfg, ctx.has_bias = csr, (heads, F_out, concat, slope), bias is not None
        ctx.shapes = (att_src.shape, att_dst.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        L = _lib.lib()
        xc, Wc, as_c, ad_c, xw, a_s, a_d, alpha = ctx.saved_tensors
        csr = ctx.csr
        heads, F_out, concat, slope = ctx.cfg
        Z, F_in = xc.shape
        HF = heads * F_out
        gc = g.contiguous().float()
        dev = gc.device
        gx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        gW = torch.empty_like(Wc)
        gas = torch.empty(HF, dtype=torch.float32, device=dev)
        gad = torch.empty(HF, dtype=torch.float32, device=dev)
        gb = torch.empty(HF if concat else F_out, dtype=torch.float32, device=dev) if ctx.has_bias else None
        nbytes = L.ab200_gat_backward_workspace_bytes(Z, csr.nnz, heads, F_out)
        ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
        rc = L.ab200_gat_backward(csr.rowptr.data_ptr(), csr.col.data_ptr(), csr.rowptr_t.data_ptr(), csr.col_t.data_ptr(),
                                  csr.eid_t.data_ptr(), Z, csr.nnz, xc.data_ptr(), F_in, Wc.data_ptr(), as_c.data_ptr(),
                                  ad_c.data_ptr(), heads, F_out, 1 if concat else 0, float(slope), xw.data_ptr(), a_s.data_ptr(),
                                  a_d.data_ptr(), alpha.data_ptr(), gc.data_ptr(), None if gx is None else gx.data_ptr(),
                                  gW.data_ptr(), gas.data_ptr(), gad.data_ptr(), None if gb is None else gb.data_ptr(),
                                  ws.data_ptr(), ws.numel(), _sp())
        _lib.check(rc, "ab200_gat_backward")
        return gx, gW, gas.view(ctx.shapes[0]), gad.view(ctx.shapes[1]), gb, None, None, None, None, None


class _Lin(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))


class GATEmbed(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True, negative_slope: float = 0.2,
                 bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.negative_slope = negative_slope
        self.lin = _Lin(in_channels, heads * out_channels)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.empty(heads * out_channels if concat else out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        for t in (self.lin.weight, self.att_src, self.att_dst):       # PyG `glorot`
            a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
            with torch.no_grad():
                t.uniform_(-a, a)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x: torch.Tensor, graph) -> torch.Tensor:
        """`graph` is a `ZoneCSR` (preferred: built once) or a PyG-style `edge_index[2, E]`."""
        csr = graph if isinstance(graph, ZoneCSR) else build_zone_csr(graph, x.shape[0]).to(x.device)
        return _GATFunction.apply(x, self.lin.weight, self.att_src, self.att_dst, self.bias, csr, self.heads, self.out_channels,
                                  self.concat, self.negative_slope)


def gnn_embed(x: torch.Tensor, edge_index: torch.Tensor, layer: GATEmbed) -> torch.Tensor:
    """Functional form: zone features [Z, F] + edge_index -> zone embedding table."""
    return layer(x, edge_index)
