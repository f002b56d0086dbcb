"""Pin oracle/augment.py against fixtures made by the reference's own utils/augment_eeg.py / utils/utils.py."""
import os

import numpy as np
import torch

from oracle import augment as A


def test_masks_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "augment_ref.npz"))
    i = 0
    while f"mask{i}" in g.files:
        c, t, u0, u1, rt, seed = (int(v) for v in g[f"mask{i}_meta"])
        prob = float(g[f"mask{i}_prob"])
        torch.manual_seed(seed)
        m = A.shape_mask((c, t), [u0, u1], prob, rt)
        ref = np.unpackbits(g[f"mask{i}"], axis=1)[:, :t]
        assert m.shape == (c, t)
        assert np.array_equal(m.numpy().astype(np.uint8), ref), i
        if rt == 2:
            assert (m == m[:1]).all()
        if rt == 3:
            assert (m == m[:, :1]).all()
        i += 1
    assert i == 6


def test_shift_and_noise(golden_dir):
    g = np.load(os.path.join(golden_dir, "augment_ref.npz"))
    assert np.array_equal(A.shift_data(g["shift_in"], 4), g["shift_out"])
    np.random.seed(11)
    out = A.add_gaussian_noise(g["noise_in"], (20, 50))
    assert np.array_equal(out, g["noise_out"])            # same numpy RNG stream -> bit-exact, incl. the 2x quirk
    resid = out - 2 * g["noise_in"]
    assert 0 < np.abs(resid).mean() < 0.1


def test_plan_equals_direct_application():
    """draw_plan + apply_plan == the reference order of operations applied directly (mask -> taylor -> shift -> pad)."""
    cfg = {"noise": {"prob": 0.0, "min_snr_dB": 20, "max_snr_dB": 50},
           "mask": {"prob": 1.0, "kwargs": {"unit": [1, 40], "mask_prob": 0.25, "random_type": 1}},
           "taylor": {"prob": 1.0}, "shift": {"prob": 1.0}}
    rng = np.random.RandomState(0)
    x = rng.randn(16, 1234).astype(np.float32)
    torch.manual_seed(5); np.random.seed(5)
    plan = A.draw_plan(x.shape, cfg, max_length=6000)
    y = A.apply_plan(x, plan, 6000)
    torch.manual_seed(5); np.random.seed(5)
    assert torch.rand(1).item() >= 0.0           # noise draw
    assert torch.rand(1).item() < 1.0            # mask draw
    m = A.shape_mask(x.shape, [1, 40], 0.25, 1).numpy()
    z = x * m
    torch.rand(1)
    n0 = np.random.randint(1, 10); n1 = np.random.randint(1, 10)
    z[:, :n0] = 0; z[:, -n1:] = 0
    torch.rand(1)
    s = np.random.randint(int(6000 - 1234 - 100))
    z = A.padding_sample(A.shift_data(z, s), 6000)
    assert y.shape == (16, 6000) and np.array_equal(y, z)
    assert plan.shift == s and plan.edge0 == n0 and plan.edge1 == n1
