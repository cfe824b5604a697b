"""FusedAdam (ab200_grad_sumsq + ab200_adam_step) against clip_grad_norm_ + torch.optim.Adam, the reference's optimiser
step (mode_sep/train/train.py:68,163-164)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("weight_decay,max_norm", [(0.0, None), (1e-4, 1.0), (1e-2, 0.05)])
def test_fused_adam_matches_torch_adam_with_clipping(weight_decay, max_norm):
    import ananke_abm_b200 as ab
    dev = _cuda()
    torch.manual_seed(0)
    net_a = torch.nn.Sequential(torch.nn.Linear(18, 128), torch.nn.ReLU(), torch.nn.Linear(128, 32), torch.nn.ReLU(),
                                torch.nn.Linear(32, 7)).to(dev)
    net_b = copy.deepcopy(net_a)
    opt_a = ab.FusedAdam(net_a.parameters(), lr=3e-3, weight_decay=weight_decay, max_grad_norm=max_norm)
    opt_b = torch.optim.Adam(net_b.parameters(), lr=3e-3, weight_decay=weight_decay)
    sd_keys = list(net_a.state_dict().keys())
    g = torch.Generator().manual_seed(1)
    for it in range(12):
        x = torch.randn(64, 18, generator=g).to(dev)
        y = torch.randn(64, 7, generator=g).to(dev) * (5.0 if it % 3 == 0 else 0.1)     # some steps clip, some do not
        for net, opt in ((net_a, opt_a), (net_b, opt_b)):
            opt.zero_grad()
            (net(x) - y).square().mean().backward()
        norm_a = opt_a.step()
        norm_b = torch.nn.utils.clip_grad_norm_(net_b.parameters(), max_norm) if max_norm is not None else \
            torch.linalg.vector_norm(torch.stack([p.grad.norm() for p in net_b.parameters()]))
        opt_b.step()
        assert abs(float(norm_a) - float(norm_b)) <= 1e-5 * float(norm_b)
    for pa, pb in zip(net_a.parameters(), net_b.parameters()):
        assert float((pa - pb).abs().max()) <= 2e-6 * float(pb.abs().max()) + 1e-7
    assert list(net_a.state_dict().keys()) == sd_keys                        # checkpoints keep their layout
    assert all(p.data_ptr() >= opt_a.flat.data_ptr() for p in net_a.parameters())   # parameters are views of the flat buffer
