#!/usr/bin/env python
"""Small end-to-end pass over every CUDA entry point, meant to run under compute-sanitizer (GPU box):

    compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python tools/sanitize_run.py [parts...]

parts: render (forward + backward, SH and precomputed-colour paths, multi-batch tile lists), integrate, tetmesh, loss,
params, filter.  Sizes are tiny on purpose (the tools slow kernels down 100-1000x).  Exit code 0 = every call returned."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gaussian-opacity-fields_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import _util  # noqa: E402
import gof_synth  # noqa: E402
from diff_gaussian_rasterization import _C  # noqa: E402


def render(dev):
    # (a) ordinary scene, image size not a tile multiple; (b) big splats: tile lists longer than one 256-entry batch
    for cfg, view, sigma in ((dict(P=3000, width=136, height=88, seed=11), 3, 2.0), (dict(P=1500, width=64, height=48, seed=12, sigma_px=14.0), 5, 14.0)):
        cam, gs = gof_synth.make_scene(cfg, view=view)
        fa = _util.fwd_args(cam, gs, dev, bg=(0.2, 0.1, 0.3))
        R, color, radii, geom, binning, img = _C.rasterize_gaussians(*fa)
        grad = torch.randn(9, cam.image_height, cam.image_width, device=dev)
        _C.rasterize_gaussians_backward(*_util.bwd_args(fa, radii, geom, R, binning, img, grad))
        _C.rasterize_gaussians_backward(*_util.bwd_args(fa, radii, geom, R, binning, img, grad))     # twice on the same buffers
        torch.cuda.synchronize()
        print("render", cfg, "R =", R, "visible =", int((radii > 0).sum()), flush=True)
    cols = torch.rand(3000, 3)
    cam, gs = gof_synth.make_scene(dict(P=3000, width=136, height=88, seed=11), view=3)
    fa = _util.fwd_args(cam, gs, dev, colors_precomp=cols)
    R, color, radii, geom, binning, img = _C.rasterize_gaussians(*fa)
    _C.rasterize_gaussians_backward(*_util.bwd_args(fa, radii, geom, R, binning, img, torch.randn(9, 88, 136, device=dev)))
    _C.mark_visible(gs["means3D"].to(dev), cam.world_view_transform.to(dev), cam.full_proj_transform.to(dev))
    torch.cuda.synchronize()


def integrate(dev):
    cam, gs = gof_synth.make_scene(dict(P=3000, width=136, height=88, seed=11), view=3)
    fa = _util.fwd_args(cam, gs, dev)
    pts = ((torch.rand(20_000, 3) * 2 - 1) * 1.6).to(dev)
    out = _C.integrate_gaussians_to_points(fa[0], pts, *fa[1:])
    torch.cuda.synchronize()
    print("integrate R =", out[0], "alpha mean", float(out[2].mean()), flush=True)


def tetmesh(dev):
    import gof_tetmesh
    z = np.load(os.path.join(ROOT, "tests", "golden", "tetmesh_noisy.npz"))
    t = lambda k: torch.from_numpy(z[k]).to(dev)
    (pos, esdf), esc, faces, iv = gof_tetmesh._unbatched_marching_tetrahedra(t("vertices"), t("tets"), t("sdf"), t("scales"))
    torch.cuda.synchronize()
    print("tetmesh faces", tuple(faces.shape), flush=True)
    (pos, esdf), esc, faces, iv = gof_tetmesh._unbatched_marching_tetrahedra(t("vertices"), t("tets"), t("sdf"), t("scales"), chunk_tets=1000)
    torch.cuda.synchronize()


def loss(dev):
    import gof_loss
    cam, _ = gof_synth.make_scene(dict(P=10, width=100, height=70, seed=1), view=4)
    img = torch.rand(9, 70, 100, device=dev, requires_grad=True)
    gt = torch.rand(3, 70, 100, device=dev)
    l, _terms = gof_loss.view_loss(img, gt, cam.world_view_transform, cam.tanfovx, cam.tanfovy, 0.2, 0.05, 100.0)
    l.backward()
    torch.cuda.synchronize()
    print("loss", float(l), flush=True)


def params(dev):
    import gof_params
    P = 2001
    raw = [torch.randn(P, 3), torch.randn(P, 4), torch.randn(P, 1), torch.rand(P, 1) * 0.01, torch.randn(P, 1, 3), torch.randn(P, 15, 3)]
    raw = [r.to(dev).requires_grad_(i != 3) for i, r in enumerate(raw)]
    outs = gof_params.activate(*raw)
    sum(o.sum() for o in outs).backward()
    p = torch.randn(P * 3, device=dev)
    gof_params.adam_step(p, torch.zeros_like(p), torch.zeros_like(p), torch.randn_like(p), 1e-3, 1)
    torch.cuda.synchronize()
    print("params ok", flush=True)


def filt(dev):
    import gof_params
    z = np.load(os.path.join(ROOT, "tests", "golden", "filter3d_a.npz"))
    out = gof_params.compute_3d_filter(torch.from_numpy(z["xyz"]).to(dev), torch.from_numpy(z["cams"]).to(dev), float(z["cams"][:, 12].max()))
    torch.cuda.synchronize()
    print("filter mean", float(out.mean()), flush=True)


PARTS = {"render": render, "integrate": integrate, "tetmesh": tetmesh, "loss": loss, "params": params, "filter": filt}

if __name__ == "__main__":
    dev = torch.device("cuda")
    for name in (sys.argv[1:] or list(PARTS)):
        PARTS[name](dev)
    print("SANITIZE_RUN_DONE", flush=True)
