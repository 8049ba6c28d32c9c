#!/usr/bin/env python
"""Golden vectors for compute_3D_filter, generated HERE by exec'ing the reference's method text
(/root/reference/scene/gaussian_model.py:262-311) inside a stub class.  Output: tests/golden/filter3d_a.npz."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "gaussian-opacity-fields_b200"))
import gof_synth  # noqa: E402

src = open("/root/reference/scene/gaussian_model.py").read()
seg = src[src.index("    @torch.no_grad()\n    def compute_3D_filter(self, cameras):"):src.index("    def oneupSHdegree(self)")]
ns = {"torch": torch}
exec("class Stub:\n" + seg, ns)
stub = ns["Stub"]()
g = torch.Generator().manual_seed(3)
P = 5000
xyz = (torch.rand(P, 3, generator=g) * 2 - 1) * 3.0
xyz[:50] += 40.0                                                   # never seen by any camera
stub.get_xyz = xyz
cams, packed = [], []
for v, (W, H) in enumerate([(640, 480), (800, 600), (320, 200), (640, 480), (1280, 720)]):
    c = gof_synth.make_camera(W, H, view=7 * v + 1, radius=3.5 + 0.3 * v)
    w2v = c.world_view_transform.t()                                # [R^T | T] with R = camera.R as stored by the reference
    R = w2v[:3, :3].t().contiguous().numpy()                        # camera.R: "stored transposed"
    T = w2v[:3, 3].contiguous().numpy()
    fx, fy = W / (2 * c.tanfovx), H / (2 * c.tanfovy)
    cams.append(types.SimpleNamespace(R=R, T=T, focal_x=fx, focal_y=fy, image_width=W, image_height=H))
    packed.append(np.concatenate([R.reshape(-1), T, [fx, fy, W, H]]).astype(np.float32))
stub.compute_3D_filter(cams)
np.savez_compressed(os.path.join(HERE, "filter3d_a.npz"), xyz=xyz.numpy(), cams=np.stack(packed), filter_3D=stub.filter_3D.numpy())
print("filter range", float(stub.filter_3D.min()), float(stub.filter_3D.max()), "unseen", int((stub.filter_3D == stub.filter_3D.max()).sum()))
