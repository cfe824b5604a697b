"""ORACLE (test infrastructure) -- graph attention layer, PyG `GATConv` semantics (torch-geometric 2.6.1).

PARITY UNPINNED: the reference declares `torch-geometric>=2.6.1` (/root/reference/pyproject.toml:25, locked
/root/reference/uv.lock:2861-2862) but never imports it -- there is no GAT call site, test or golden vector in the
reference; the package is not installed here.  This file restates the published
algorithm (torch_geometric/nn/conv/gat_conv.py: lin -> alpha_src/alpha_dst -> leaky_relu -> softmax over the
incoming edges of each target -> sum, concat or mean over heads, + bias; add_self_loops=True) in two independent
ways that are checked against each other (tests/test_gat.py):
  * `gat_dense`  : dense masked-softmax over a [Z, Z] adjacency
  * `gat_edges`  : edge-list scatter formulation (the message-passing form PyG uses)
Graph inputs follow the reference's producers: `edge_index[2, E]` lists each undirected edge once
(/root/reference/src/ananke_abm/data_generator/mock_2p.py:228-230), symmetrised with self loops as
/root/reference/src/ananke_abm/data_generator/load_data.py:104-110 does.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch
import torch.nn.functional as F


def symmetrise_with_self_loops(edge_index: torch.Tensor, Z: int) -> torch.Tensor:
    """-> directed edge list [2, nnz] (source, target): both directions of every edge + one self loop per node,
    duplicates removed, sorted by (target, source)."""
    src, dst = edge_index[0].long(), edge_index[1].long()
    keep = src != dst
    src, dst = src[keep], dst[keep]
    loops = torch.arange(Z)
    s = torch.cat([src, dst, loops])
    d = torch.cat([dst, src, loops])
    key = torch.unique(d * Z + s)        # sorted by target then source
    return torch.stack([key % Z, key // Z])


def glorot_(t: torch.Tensor, gen=None) -> torch.Tensor:
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a, generator=gen)
    return t


def gat_edges(x, edges, W, att_src, att_dst, bias, heads: int, F_out: int, concat: bool = True, slope: float = 0.2):
    Z = x.shape[0]
    xw = (x @ W.T).view(Z, heads, F_out)
    a_s = (xw * att_src.view(1, heads, F_out)).sum(-1)
    a_d = (xw * att_dst.view(1, heads, F_out)).sum(-1)
    j, i = edges[0], edges[1]
    e = F.leaky_relu(a_s[j] + a_d[i], slope)                       # [nnz, H]
    m = torch.full((Z, heads), -float("inf"), dtype=x.dtype, device=x.device).scatter_reduce(0, i.view(-1, 1).expand(-1, heads), e, "amax")
    w = torch.exp(e - m[i])
    den = torch.zeros(Z, heads, dtype=x.dtype, device=x.device).index_add(0, i, w)
    alpha = w / den[i]
    out = torch.zeros(Z, heads, F_out, dtype=x.dtype, device=x.device).index_add(0, i, alpha.unsqueeze(-1) * xw[j])
    out = out.reshape(Z, heads * F_out) if concat else out.mean(dim=1)
    return out + bias if bias is not None else out


def gat_dense(x, edges, W, att_src, att_dst, bias, heads: int, F_out: int, concat: bool = True, slope: float = 0.2):
    Z = x.shape[0]
    xw = (x @ W.T).view(Z, heads, F_out)
    a_s = (xw * att_src.view(1, heads, F_out)).sum(-1)
    a_d = (xw * att_dst.view(1, heads, F_out)).sum(-1)
    adj = torch.zeros(Z, Z, dtype=torch.bool)
    adj[edges[1], edges[0]] = True                                  # adj[i, j]: j -> i
    e = F.leaky_relu(a_d.unsqueeze(1) + a_s.unsqueeze(0), slope)    # [i, j, H]
    e = e.masked_fill(~adj.unsqueeze(-1), -float("inf"))
    alpha = torch.softmax(e, dim=1)
    out = torch.einsum("ijh,jhf->ihf", alpha, xw)
    out = out.reshape(Z, heads * F_out) if concat else out.mean(dim=1)
    return out + bias if bias is not None else out


def synthetic_zone_graph(Z: int, k: int = 6, seed: int = 42) -> Tuple[torch.Tensor, torch.Tensor]:
    """jittered sqrt(Z) x sqrt(Z) lattice, undirected k-NN graph; returns (edge_index [2,E] with
    each undirected edge once, zone features [Z,7] ~ U[0,1))."""
    g = torch.Generator().manual_seed(seed)
    side = int(math.ceil(math.sqrt(Z)))
    idx = torch.arange(Z)
    xy = torch.stack([(idx % side).float(), (idx // side).float()], dim=-1) + torch.rand(Z, 2, generator=g)
    feats = torch.rand(Z, 7, generator=g)
    # k nearest neighbours via a cell grid: candidates within +-2 lattice cells
    cand = []
    for dx in range(-2, 3):
        for dy in range(-2, 3):
            if dx == 0 and dy == 0:
                continue
            cx, cy = idx % side + dx, idx // side + dy
            ok = (cx >= 0) & (cx < side) & (cy >= 0) & (cy < side)
            nb = cy * side + cx
            ok = ok & (nb < Z)
            cand.append(torch.where(ok, nb, torch.full_like(nb, -1)))
    cand = torch.stack(cand, dim=1)                                  # [Z, 24]
    d = (xy.unsqueeze(1) - xy[cand.clamp(min=0)]).pow(2).sum(-1)
    d = torch.where(cand >= 0, d, torch.full_like(d, float("inf")))
    nn_idx = d.topk(k, dim=1, largest=False).indices
    nbr = cand.gather(1, nn_idx)
    src = idx.unsqueeze(1).expand(-1, k).reshape(-1)
    dst = nbr.reshape(-1)
    ok = dst >= 0
    src, dst = src[ok], dst[ok]
    lo, hi = torch.minimum(src, dst), torch.maximum(src, dst)
    key = torch.unique(lo * Z + hi)
    return torch.stack([key // Z, key % Z]), feats
