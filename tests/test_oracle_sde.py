"""The SDE oracle (oracle/sde_oracle.py): Philox4x32-10 pinned on the Random123 known-answer vectors, the normals'
moments, and the Euler-Maruyama stepping rule on a linear SDE with known mean."""
import numpy as np
import torch

from oracle import sde_oracle as so


def test_philox4x32_10_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = so.philox4x32_10(*ctr, *key)
        assert tuple(int(v) for v in got) == want


def test_normals_moments_and_counter_independence():
    z = so.normals(4096, 64, seed=11, step=3)
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01
    assert abs(float((z ** 4).mean()) - 3.0) < 0.1
    # counter-based: the first rows of a bigger batch are the smaller batch's rows; another step / seed differs
    assert np.array_equal(so.normals(16, 64, 11, 3), z[:16])
    assert not np.array_equal(so.normals(16, 64, 11, 4), z[:16]) and not np.array_equal(so.normals(16, 64, 12, 3), z[:16])


def test_euler_maruyama_rule_on_linear_sde():
    class Lin:
        def f(self, t, y):
            return -0.5 * y

        def g(self, t, y):
            return torch.full_like(y, 0.2)

    y0 = torch.ones(20000, 4)
    ts = torch.tensor([0.0, 0.25, 0.505, 1.0])
    path = so.sdeint_euler(Lin(), y0, ts, dt=0.01, seed=1)
    assert path.shape == (4, 20000, 4) and torch.equal(path[0], y0)
    mean = path.mean(dim=(1, 2))
    want = (1 - 0.5 * 0.01) ** torch.tensor([0.0, 25.0, 50.5, 100.0])       # Euler mean; 0.505 is read off by interpolation
    assert float((mean - want).abs().max()) < 5e-3
    # sigma = 0 reduces to explicit Euler exactly
    class Det(Lin):
        def g(self, t, y):
            return torch.zeros_like(y)
    p2 = so.sdeint_euler(Det(), torch.ones(2, 4), torch.tensor([0.0, 0.03]), dt=0.01, seed=1)
    assert torch.allclose(p2[1], torch.full((2, 4), (1 - 0.005) ** 3), atol=1e-6)
