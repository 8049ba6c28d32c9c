#!/usr/bin/env python
"""One integrate call for ncu: python tools/one_integrate.py C3 2000000 [ours|ref]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _util, gof_synth
from diff_gaussian_rasterization import _C as ours
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
n_pts = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
which = sys.argv[3] if len(sys.argv) > 3 else "ours"
mod = ours if which == "ours" else _util.load_ref()
dev = torch.device("cuda")
cam, gs = gof_synth.make_scene(name, view=3)
P = gs["means3D"].shape[0]
g = torch.Generator().manual_seed(4)
idx = torch.randint(0, P, (n_pts,), generator=g)
pts = (gs["means3D"][idx] + gs["scales"][idx] * 3.0 * (torch.rand(n_pts, 3, generator=g) * 2 - 1)).contiguous().to(dev)
fa = _util.fwd_args(cam, gs, dev)
out = mod.integrate_gaussians_to_points(*((fa[0], pts) + tuple(fa[1:])))
torch.cuda.synchronize()
print("done", out[0])
