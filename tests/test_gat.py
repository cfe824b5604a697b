"""GAT: the two oracle formulations against each other (CPU), CSR construction (CPU), CUDA kernels vs oracle (GPU)."""
import pytest
import torch

from oracle import gat_oracle as go


def _params(F_in, heads, F_out, concat, seed=0):
    g = torch.Generator().manual_seed(seed)
    W = go.glorot_(torch.empty(heads * F_out, F_in), g)
    a_s = go.glorot_(torch.empty(1, heads, F_out), g)
    a_d = go.glorot_(torch.empty(1, heads, F_out), g)
    b = torch.randn(heads * F_out if concat else F_out, generator=g) * 0.1
    return W, a_s, a_d, b


MOCK_EDGES = torch.tensor([[0, 0, 0, 1, 1, 1, 2, 2, 3, 3, 4, 6], [1, 6, 7, 2, 5, 6, 3, 5, 4, 6, 5, 7]])


@pytest.mark.parametrize("heads,F_out,concat", [(1, 8, True), (4, 16, True), (4, 16, False)])
def test_oracle_dense_equals_edge_formulation(heads, F_out, concat):
    Z = 8
    x = torch.rand(Z, 7, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    W, a_s, a_d, b = (t.double() for t in _params(7, heads, F_out, concat))
    edges = go.symmetrise_with_self_loops(MOCK_EDGES, Z)
    assert edges.shape[1] == 2 * 12 + 8
    o1 = go.gat_edges(x, edges, W, a_s, a_d, b, heads, F_out, concat)
    o2 = go.gat_dense(x, edges, W, a_s, a_d, b, heads, F_out, concat)
    assert torch.allclose(o1, o2, atol=1e-12)
    # attention rows sum to one: with W=0 and bias=0 the output is 0; with identical xw rows the output is that row
    xc = torch.ones(Z, 7, dtype=torch.float64)
    o3 = go.gat_edges(xc, edges, W, a_s, a_d, None, heads, F_out, True)
    assert torch.allclose(o3, (xc @ W.T), atol=1e-12)


def test_zone_csr_matches_oracle_edge_order():
    from ananke_abm_b200.graph import build_zone_csr
    csr = build_zone_csr(MOCK_EDGES, 8)
    edges = go.symmetrise_with_self_loops(MOCK_EDGES, 8)
    assert csr.nnz == 32 and torch.equal(csr.edges, edges)
    assert int(csr.rowptr[-1]) == 32 and int(csr.rowptr_t[-1]) == 32
    for i in range(8):                       # every row holds its self loop and is sorted by source
        cols = csr.col[csr.rowptr[i]:csr.rowptr[i + 1]].tolist()
        assert i in cols and cols == sorted(cols)
    # the transposed CSR enumerates the same edges
    for j in range(8):
        for e in range(int(csr.rowptr_t[j]), int(csr.rowptr_t[j + 1])):
            eid = int(csr.eid_t[e])
            assert int(csr.col[eid]) == j and int(csr.edges[1, eid]) == int(csr.col_t[e])
    # ragged / degenerate inputs: isolated node keeps only its self loop; duplicate and reversed edges collapse
    c2 = build_zone_csr(torch.tensor([[0, 1, 1], [1, 0, 1]]), 3)
    assert c2.nnz == 5 and c2.col[c2.rowptr[2]:c2.rowptr[3]].tolist() == [2]


def test_synthetic_graph_degree():
    ei, feats = go.synthetic_zone_graph(500, k=6)
    from ananke_abm_b200.graph import build_zone_csr, synthetic_zone_graph
    ei2, feats2 = synthetic_zone_graph(500, k=6)
    assert torch.equal(ei, ei2) and torch.equal(feats, feats2)
    csr = build_zone_csr(ei, 500)
    deg = (csr.rowptr[1:] - csr.rowptr[:-1]).float()
    assert feats.shape == (500, 7) and 6.5 < float(deg.mean()) < 10.0 and int(deg.min()) >= 2


@pytest.mark.gpu
@pytest.mark.parametrize("Z,heads,F_out,concat", [(8, 1, 8, True), (8, 4, 16, True), (500, 4, 16, True), (500, 4, 16, False),
                                                   (10000, 4, 16, True), (37, 2, 32, True)])
def test_gat_cuda_forward_backward_vs_oracle(Z, heads, F_out, concat):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from ananke_abm_b200.gnn_embed import GATEmbed
    from ananke_abm_b200.graph import build_zone_csr
    dev = torch.device("cuda:0")
    if Z == 8:
        ei, x = MOCK_EDGES, torch.rand(8, 7, generator=torch.Generator().manual_seed(2))
    else:
        ei, x = go.synthetic_zone_graph(Z, k=6, seed=Z)
    W, a_s, a_d, b = _params(7, heads, F_out, concat, seed=Z)
    edges = go.symmetrise_with_self_loops(ei, Z)
    wgt = torch.randn(Z, heads * F_out if concat else F_out, generator=torch.Generator().manual_seed(3))

    ref_in = [t.clone().requires_grad_(True) for t in (x, W, a_s, a_d, b)]
    ref = go.gat_edges(ref_in[0], edges, ref_in[1], ref_in[2], ref_in[3], ref_in[4], heads, F_out, concat)
    (ref * wgt).sum().backward()

    layer = GATEmbed(7, F_out, heads=heads, concat=concat).to(dev)
    with torch.no_grad():
        layer.lin.weight.copy_(W); layer.att_src.copy_(a_s); layer.att_dst.copy_(a_d); layer.bias.copy_(b)
    csr = build_zone_csr(ei, Z).to(dev)
    xd = x.to(dev).requires_grad_(True)
    out = layer(xd, csr)
    (out * wgt.to(dev)).sum().backward()

    def rel(a, b_):
        return float((a.double().cpu() - b_.double()).abs().max() / b_.double().abs().max().clamp_min(1e-30))
    assert rel(out.detach(), ref.detach()) < 1e-5
    assert rel(xd.grad, ref_in[0].grad) < 2e-5
    assert rel(layer.lin.weight.grad, ref_in[1].grad) < 2e-5
    assert rel(layer.att_src.grad, ref_in[2].grad) < 2e-5
    assert rel(layer.att_dst.grad, ref_in[3].grad) < 2e-5
    assert rel(layer.bias.grad, ref_in[4].grad) < 2e-5
