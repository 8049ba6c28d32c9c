#!/usr/bin/env python
"""One fwd+bwd of a config for ncu launch lists: python tools/one_iter.py C3 [ours|ref] [iters]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _util, gof_synth
from diff_gaussian_rasterization import _C as ours
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
which = sys.argv[2] if len(sys.argv) > 2 else "ours"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 1
mod = ours if which == "ours" else _util.load_ref()
dev = torch.device("cuda")
cam, gs = gof_synth.make_scene(name, view=1)
fa = _util.fwd_args(cam, gs, dev)
grad = torch.randn(9, cam.image_height, cam.image_width, device=dev)
torch.cuda.synchronize()
for _ in range(iters):
    R, color, radii, geom, binning, img = mod.rasterize_gaussians(*fa)
    g = mod.rasterize_gaussians_backward(*_util.bwd_args(fa, radii, geom, R, binning, img, grad))
torch.cuda.synchronize()
print("done", R)
