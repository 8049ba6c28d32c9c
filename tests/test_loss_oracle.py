"""CPU: the loss oracle (oracle/loss_oracle.py) against golden vectors produced by the reference's own Python
(tests/golden/make_golden_loss.py): every term of train.py:151-188 and the gradient with respect to the render."""
import glob
import os

import numpy as np
import pytest

import loss_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = sorted(glob.glob(os.path.join(HERE, "golden", "loss_*.npz")))


def test_fixtures_present():
    assert len(FIX) >= 3


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[:-4] for p in FIX])
def test_loss_oracle_matches_reference(path):
    fx = np.load(path)
    out = loss_oracle.view_loss(fx["render"], fx["gt"], fx["world_view_transform"], float(fx["tanfovx"]), float(fx["tanfovy"]), fx["lambdas"])
    for k in ("Ll1", "ssim", "distortion_loss", "depth_normal_loss", "loss"):
        assert abs(out[k] - float(fx[k])) <= 2e-6 * max(1.0, abs(float(fx[k]))), k
    assert np.abs(out["depth_normal"] - fx["depth_normal"]).max() < 2e-5
    g, r = out["grad"], fx["grad"].astype(np.float64)
    for ch in range(9):
        den = max(np.abs(r[ch]).max(), 1e-12)
        assert np.abs(g[ch] - r[ch]).max() / den < 1e-4, f"channel {ch}"   # the goldens are float32 autograd
    assert np.all(g[7] == 0) and np.all(r[7] == 0)
