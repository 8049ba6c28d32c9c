"""GPU: size-independent properties at BASELINE's full size (config 3: 1M Gaussians, 1920x1080), through the C ABI.
Sortedness of every tile list by (depth bits, Gaussian index), tile ranges partition the instance list, forward is
bit-deterministic, alpha channel == 1 - T, backward is linear in dL/dout, n_contrib never exceeds the list length."""
import pytest
import torch

import _util
import gof_synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3():
    from diff_gaussian_rasterization import _C
    dev = torch.device("cuda")
    cam, gs = gof_synth.make_scene("C3", view=5)
    fa = _util.fwd_args(cam, gs, dev)
    out = _C.rasterize_gaussians(*fa)
    P, W, H = gs["means3D"].shape[0], cam.image_width, cam.image_height
    st = _C.export_state(P, W, H, out[0], out[3], out[4], out[5], out[2])
    return _C, fa, out, st, (P, W, H)


def test_binning_invariants(c3):
    _C, fa, (R, color, radii, geom, binning, img), st, (P, W, H) = c3
    tiles = st["tiles_touched"].long()
    assert int(tiles.sum()) == R
    assert torch.equal(tiles > 0, radii > 0)
    ranges = st["ranges"].long()
    lens = ranges[:, 1] - ranges[:, 0]
    assert int(lens.sum()) == R and int(lens.min()) >= 0
    nz = lens > 0
    starts = ranges[nz, 0]
    assert torch.equal(starts, torch.cat([torch.zeros(1, dtype=torch.long, device=starts.device), ranges[nz, 1][:-1]]))
    # each tile list sorted by (depth bits, gaussian index): build the reference's 64-bit key and check monotonicity
    pl = st["point_list"].long()
    tile_of = torch.repeat_interleave(torch.arange(ranges.shape[0], device=pl.device), lens)
    depth_bits = st["depths"].view(torch.int32).long()[pl]
    key = (tile_of << 32) | depth_bits
    assert bool((key[1:] >= key[:-1]).all())
    same = key[1:] == key[:-1]
    assert bool((pl[1:][same] > pl[:-1][same]).all()), "ties must keep ascending Gaussian index (stable sort)"


def test_forward_is_deterministic_and_consistent(c3):
    _C, fa, (R, color, radii, geom, binning, img), st, (P, W, H) = c3
    R2, color2, radii2, *_ = _C.rasterize_gaussians(*fa)
    assert R2 == R and torch.equal(color, color2) and torch.equal(radii, radii2)
    T = st["accum_alpha"][0]
    assert torch.allclose(color[7], 1.0 - T, atol=2e-5)
    lens = (st["ranges"][:, 1] - st["ranges"][:, 0]).long()
    ty = torch.arange(H, device=color.device) // 16
    tx = torch.arange(W, device=color.device) // 16
    per_pixel_len = lens[(ty[:, None] * ((W + 15) // 16) + tx[None, :])]
    assert bool((st["n_contrib"][0].long() <= per_pixel_len).all())
    assert float(color[:3].min()) >= 0.0 and torch.isfinite(color).all()


def test_backward_is_linear_in_upstream_gradient(c3):
    _C, fa, (R, color, radii, geom, binning, img), st, (P, W, H) = c3
    dev = color.device
    g = torch.randn(9, H, W, generator=torch.Generator().manual_seed(1)).to(dev)
    a = _C.rasterize_gaussians_backward(*_util.bwd_args(fa, radii, geom, R, binning, img, g))
    b = _C.rasterize_gaussians_backward(*_util.bwd_args(fa, radii, geom, R, binning, img, 2.0 * g))
    for name, x, y in zip(["dmeans2D", "dcolors", "dopacity", "dsh", "dv2g"], (a[0], a[1], a[2], a[5], a[8]), (b[0], b[1], b[2], b[5], b[8])):
        if name == "dmeans2D":   # the third column is an abs-sum statistic: also linear for a positive factor
            pass
        assert _util.rel_err(y, 2.0 * x)[0] < 1e-4, name
    zero = _C.rasterize_gaussians_backward(*_util.bwd_args(fa, radii, geom, R, binning, img, torch.zeros_like(g)))
    assert all(float(t.abs().max()) == 0.0 for t in zero)
