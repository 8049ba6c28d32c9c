"""GPU: the fused per-view loss (gof_loss.view_loss -> gof_view_loss, csrc/view_loss.cu) against the golden vectors the
reference's own Python produced (tests/golden/make_golden_loss.py) and, at 1080p, against the oracle on a crop-free
statistic: the same kernel source is checked phase by phase on the CPU in test_view_loss_host.py."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FIX = sorted(glob.glob(os.path.join(HERE, "golden", "loss_*.npz")))


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[:-4] for p in FIX])
def test_view_loss_matches_reference_goldens(path):
    import gof_loss
    fx = np.load(path)
    dev = torch.device("cuda")
    rendering = torch.from_numpy(fx["render"]).to(dev).requires_grad_(True)
    gt = torch.from_numpy(fx["gt"]).to(dev)
    lam = [float(x) for x in fx["lambdas"]]
    loss, terms = gof_loss.view_loss(rendering, gt, torch.from_numpy(fx["world_view_transform"]), float(fx["tanfovx"]),
                                     float(fx["tanfovy"]), lam[0], lam[1], lam[2])
    (2.0 * loss).backward()
    t = terms.cpu().numpy()
    for i, k in enumerate(("Ll1", "ssim", "depth_normal_loss", "distortion_loss", "loss")):
        assert abs(float(t[i]) - float(fx[k])) <= 1e-5 * max(1.0, abs(float(fx[k]))), k
    assert abs(float(loss.detach()) - float(fx["loss"])) <= 1e-5 * max(1.0, abs(float(fx["loss"])))
    g, r = rendering.grad.cpu().numpy() / 2.0, fx["grad"].astype(np.float64)
    for ch in range(9):
        den = max(np.abs(r[ch]).max(), 1e-12)
        assert np.abs(g[ch] - r[ch]).max() / den < 1e-4, f"channel {ch}"


def test_view_loss_1080p_runs_and_is_deterministic():
    import gof_loss
    import gof_synth
    dev = torch.device("cuda")
    cam = gof_synth.make_camera(1920, 1080, view=7)
    g = torch.Generator().manual_seed(5)
    rendering = torch.rand(9, 1080, 1920, generator=g).to(dev).requires_grad_(True)
    gt = torch.rand(3, 1080, 1920, generator=g).to(dev)
    rot = gof_loss.camera_rotation(cam.world_view_transform)
    out = []
    for _ in range(2):
        rendering.grad = None
        loss, terms = gof_loss.view_loss(rendering, gt, cam.world_view_transform, cam.tanfovx, cam.tanfovy, rotation=rot)
        loss.backward()
        out.append((terms.cpu().numpy().copy(), rendering.grad.cpu().numpy().copy()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert np.isfinite(out[0][0]).all() and np.isfinite(out[0][1]).all()
    # mean |rgb - gt| of two independent uniforms is 1/3; distortion channel mean 1/2
    assert abs(out[0][0][0] - 1.0 / 3.0) < 2e-3 and abs(out[0][0][3] - 0.5) < 2e-3
