"""GPU: the one-sweep binning (csrc/binning.cu: single-launch radix passes with decoupled look-back, fused scan + instance
emission, num_rendered summed by the preprocess kernel and read on a side stream) against the round-1 multi-launch kernels
(csrc/binning_legacy.cu) -- every index buffer bit-exact, over sizes that exercise one / two tile digits, ragged tails, chunks
with empty digits, a single chunk, and tile lists of big splats.  (Both are compared with the live reference elsewhere.)"""
import pytest
import torch

import _util
import gof_synth

pytestmark = pytest.mark.gpu

CASES = [
    dict(P=1, width=16, height=16, seed=1),                       # one Gaussian, one tile
    dict(P=257, width=64, height=48, seed=2),                     # a few chunks' worth of nothing: single-chunk sorts
    dict(P=5_000, width=256, height=256, seed=3),                 # 256 tiles: ONE tile digit
    dict(P=4_097, width=272, height=256, seed=4),                 # 272 tiles: two digits, P just over one sort chunk
    dict(P=70_001, width=800, height=608, seed=5),
    dict(P=3_000, width=320, height=240, seed=6, sigma_px=25.0),  # big splats: thousands of instances per CTA of the emit kernel
    dict(P=300_000, width=1920, height=1080, seed=7),
]


@pytest.mark.parametrize("cfg", CASES, ids=[f"P{c['P']}_{c['width']}x{c['height']}" for c in CASES])
def test_onesweep_binning_equals_legacy(cfg):
    from diff_gaussian_rasterization import _C
    dev = torch.device("cuda")
    cam, gs = gof_synth.make_scene(cfg, view=cfg["seed"])
    fa = _util.fwd_args(cam, gs, dev)
    P, W, H = cfg["P"], cfg["width"], cfg["height"]
    out = {}
    try:
        for mode in (1, 0, 0):                                   # legacy, one-sweep, one-sweep again (reused scratch buffers)
            _C._lib.gof_set_binning_legacy(mode)
            R, color, radii, geom, binning, img = _C.rasterize_gaussians(*fa)
            st = _C.export_state(P, W, H, R, geom, binning, img, radii)
            torch.cuda.synchronize()
            cur = (R, radii.clone(), st["point_list"].clone(), st["ranges"].clone(), st["n_contrib"].clone(), color.clone())
            if mode == 1:
                out["legacy"] = cur
            else:
                a, b = out["legacy"], cur
                assert a[0] == b[0], "num_rendered"
                assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])
                assert torch.equal(a[5], b[5]), "same lists -> same image bits"
    finally:
        _C._lib.gof_set_binning_legacy(0)


def test_sorts_and_scan_of_the_extraction_path_equal_legacy():
    """integrate (point sort by tile) and marching tetrahedra (pair sorts + scans) through both binning implementations."""
    import numpy as np
    import os
    import gof_tetmesh
    from diff_gaussian_rasterization import _C
    dev = torch.device("cuda")
    cam, gs = gof_synth.make_scene(dict(P=20_000, width=400, height=304, seed=9), view=3)
    fa = _util.fwd_args(cam, gs, dev)
    pts = ((torch.rand(300_007, 3, generator=torch.Generator().manual_seed(1)) * 2 - 1) * 1.6).to(dev)
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tetmesh_noisy.npz"))
    t = lambda k: torch.from_numpy(z[k]).to(dev)
    res = []
    try:
        for mode in (1, 0):
            _C._lib.gof_set_binning_legacy(mode)
            o = _C.integrate_gaussians_to_points(fa[0], pts, *fa[1:])
            (pos, esdf), esc, faces, iv = gof_tetmesh._unbatched_marching_tetrahedra(t("vertices"), t("tets"), t("sdf"), t("scales"))
            torch.cuda.synchronize()
            res.append((o[0], o[1].clone(), o[2].clone(), o[3].clone(), faces.clone(), iv.clone(), pos.clone()))
    finally:
        _C._lib.gof_set_binning_legacy(0)
    a, b = res
    assert a[0] == b[0]
    for x, y in zip(a[1:], b[1:]):
        assert torch.equal(x, y)
