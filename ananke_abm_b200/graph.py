"""Zone-graph preprocessing for the GAT kernels: symmetrise + self loops, CSR sorted by destination, and the
transposed CSR with the edge-id map the backward pass needs.  Host-side, done once per graph.

Inputs follow the reference's producers: `edge_index[2, E]` with each undirected edge listed once
(/root/reference/src/ananke_abm/data_generator/mock_2p.py:228-230); the symmetrisation + self loops match the
adjacency built at /root/reference/src/ananke_abm/data_generator/load_data.py:104-110.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Tuple

import torch


@dataclass
class ZoneCSR:
    Z: int
    nnz: int
    rowptr: torch.Tensor      # [Z+1] int32, by destination
    col: torch.Tensor         # [nnz] int32, source of each in-edge
    rowptr_t: torch.Tensor    # [Z+1] int32, by source
    col_t: torch.Tensor       # [nnz] int32, destination of each out-edge
    eid_t: torch.Tensor       # [nnz] int32, position of that edge in the by-destination order
    edges: torch.Tensor       # [2, nnz] int64 (source, target) in by-destination order

    def to(self, device) -> "ZoneCSR":
        return ZoneCSR(self.Z, self.nnz, *(t.to(device) for t in (self.rowptr, self.col, self.rowptr_t, self.col_t,
                                                                   self.eid_t, self.edges)))


def build_zone_csr(edge_index: torch.Tensor, Z: int, symmetrise: bool = True, add_self_loops: bool = True) -> ZoneCSR:
    ei = edge_index.detach().to("cpu").long()
    src, dst = ei[0], ei[1]
    if add_self_loops:
        keep = src != dst
        src, dst = src[keep], dst[keep]
    s, d = [src], [dst]
    if symmetrise:
        s.append(dst)
        d.append(src)
    if add_self_loops:
        loops = torch.arange(Z)
        s.append(loops)
        d.append(loops)
    s, d = torch.cat(s), torch.cat(d)
    key = torch.unique(d * Z + s)                 # sorted by (destination, source), duplicates removed
    d, s = key // Z, key % Z
    nnz = int(key.numel())
    rowptr = torch.zeros(Z + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(d, minlength=Z), 0)
    # transposed: sort the same edges by (source, destination)
    key_t = s * Z + d
    order = torch.argsort(key_t, stable=True)
    rowptr_t = torch.zeros(Z + 1, dtype=torch.int64)
    rowptr_t[1:] = torch.cumsum(torch.bincount(s, minlength=Z), 0)
    return ZoneCSR(Z, nnz, rowptr.int(), s.int(), rowptr_t.int(), d[order].int(), order.int(), torch.stack([s, d]))


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
