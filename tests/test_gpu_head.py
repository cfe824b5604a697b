"""GPU parity of the fused classification head + argmax (`ab200_head_argmax`) against the reference's formulation
(mode_sep/architecture/model.py:196-199, inference.py:57): identical predicted labels, logits to fp32 round-off."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _ref(pred_emb, table, tau):
    tn = table / (table.norm(dim=-1, keepdim=True) + 1e-8)
    en = pred_emb / (pred_emb.norm(dim=-1, keepdim=True) + 1e-8)
    logits = torch.einsum("me,ze->mz", en.double(), tn.double()) / tau       # fp64 ground truth
    return logits


@pytest.mark.parametrize("M,Z", [(1, 1), (5, 8), (127, 129), (128, 128), (1000, 500), (4097, 1000), (20000, 10000)])
def test_head_argmax_matches_dense_reference(M, Z):
    from ananke_abm_b200.inference import head_argmax
    dev = _cuda()
    g = torch.Generator().manual_seed(M * 7 + Z)
    emb = torch.randn(M, 64, generator=g) * torch.rand(M, 1, generator=g) * 3.0
    table = torch.randn(Z, 64, generator=g)
    labels, best = head_argmax(emb.to(dev), table.to(dev), 0.2, return_logit=True)
    torch.cuda.synchronize()
    logits = _ref(emb, table, 0.2)
    ref_lab = logits.argmax(-1)
    ref_best = logits.max(-1).values
    assert labels.dtype == torch.int64 and labels.shape == (M,)
    assert torch.equal(labels.cpu(), ref_lab)
    assert float((best.cpu().double() - ref_best).abs().max()) < 2e-5 * float(ref_best.abs().max().clamp_min(1.0))


def test_head_argmax_ties_take_the_first_index_and_zero_rows_are_safe():
    from ananke_abm_b200.inference import head_argmax
    dev = _cuda()
    table = torch.randn(40, 64, generator=torch.Generator().manual_seed(0))
    table[17] = table[5]                       # exact duplicate zone: the lower index must win
    emb = torch.cat([table[5:6] * 2.0, torch.zeros(1, 64), table[30:31]], dim=0)
    labels = head_argmax(emb.to(dev), table.to(dev), 0.2).cpu()
    assert labels.tolist()[0] == 5 and labels.tolist()[2] == 30
    assert 0 <= labels.tolist()[1] < 40        # all-zero embedding: every logit is 0, any valid zone (reference: index 0)
    assert labels.tolist()[1] == 0


def test_predict_labels_fused_equals_module_head(golden_mode_sep):
    """end to end on the reference fixture: fused head == logits.argmax of the module == golden labels"""
    import ananke_abm_b200 as ab
    dev = _cuda()
    g = golden_mode_sep
    m = ab.ModeSepModel(8, ab.ModeSepConfig())
    m.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd_")}, strict=True)
    m = m.to(dev)
    t = torch.from_numpy(g["times_union"]).to(dev)
    home, work, traits = (torch.from_numpy(g[k]).to(dev) for k in ("home_idx", "work_idx", "traits"))
    fused = ab.inference.predict_labels(m, t, home, work, traits)
    dense = ab.inference.predict_labels(m, t, home, work, traits, fused=False)
    assert torch.equal(fused, dense)
    assert np.array_equal(fused.cpu().numpy(), g["labels"])
