#!/usr/bin/env python
"""Golden vectors for the parameter prologue / epilogue around the rasterizer, generated HERE
from the reference's own Python: the property bodies `get_scaling_with_3D_filter`, `get_rotation`, `get_features`,
`get_opacity_with_3D_filter` are cut out of /root/reference/scene/gaussian_model.py as TEXT and exec'ed inside a stub
class (the module itself cannot be imported: it needs the simple_knn extension), the three activations are read from
its `setup_functions`; gradients come from autograd, the optimizer step from torch.optim.Adam with the reference's
`eps=1e-15` (gaussian_model.py:360).  Output: tests/golden/params_*.npz."""
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/scene/gaussian_model.py"


def build_stub():
    src = open(REF).read()
    seg = src[src.index("    @property\n    def get_scaling(self):"):src.index("    def get_apperance_embedding")]
    ns = {"torch": torch}
    exec("class Stub:\n" + seg, ns)
    stub = ns["Stub"]()
    setup = src[src.index("    def setup_functions(self):"):src.index("    def __init__(self, sh_degree")]
    for name, expr in re.findall(r"self\.(scaling_activation|opacity_activation|rotation_activation) = ([\w\.]+)", setup):
        setattr(stub, name, eval(expr, {"torch": torch}))
    return stub


def main():
    stub = build_stub()
    for name, P, seed in (("params_a", 257, 1), ("params_b", 64, 2)):
        g = torch.Generator().manual_seed(seed)
        raw = {
            "_xyz": torch.randn(P, 3, generator=g),
            "_scaling": torch.log(torch.rand(P, 3, generator=g) * 0.05 + 1e-3),
            "_rotation": torch.randn(P, 4, generator=g),
            "_opacity": torch.randn(P, 1, generator=g) * 2.0,
            "_features_dc": torch.randn(P, 1, 3, generator=g),
            "_features_rest": torch.randn(P, 15, 3, generator=g) * 0.1,
        }
        raw["_rotation"][0] = 0.0                                   # degenerate quaternion: F.normalize's eps branch
        filter_3D = torch.rand(P, 1, generator=g) * 0.02
        filter_3D[1] = 0.0
        for k, v in raw.items():
            setattr(stub, k, v.clone().requires_grad_(True))
        stub.filter_3D = filter_3D
        outs = {"scales": stub.get_scaling_with_3D_filter, "rotations": stub.get_rotation,
                "opacities": stub.get_opacity_with_3D_filter, "shs": stub.get_features}
        up = {k: torch.randn(v.shape, generator=g) for k, v in outs.items()}          # upstream gradients (the rasterizer's)
        sum((outs[k] * up[k]).sum() for k in outs).backward()
        save = {"filter_3D": filter_3D.numpy()}
        for k, v in raw.items():
            save["raw" + k] = v.numpy()
            save["grad" + k] = getattr(stub, k).grad.numpy() if k != "_xyz" else np.zeros_like(v.numpy())
        for k in outs:
            save["out_" + k] = outs[k].detach().numpy()
            save["up_" + k] = up[k].numpy()
        # ---- optimizer: three Adam steps on one tensor with changing gradients (gaussian_model.py:360) ----
        p0 = torch.randn(P, 7, generator=g)
        p = p0.clone().requires_grad_(True)
        opt = torch.optim.Adam([{"params": [p], "lr": 1.6e-4}], lr=0.0, eps=1e-15)
        grads = [torch.randn(P, 7, generator=g) * (10.0 ** (-i)) for i in range(3)]
        grads[1][::5] = 0.0                                          # Gaussians invisible in a view still take the step
        for gi in grads:
            p.grad = gi.clone()
            opt.step()
        st = opt.state[p]
        save.update(adam_p0=p0.numpy(), adam_grads=np.stack([x.numpy() for x in grads]), adam_p=p.detach().numpy(),
                    adam_m=st["exp_avg"].numpy(), adam_v=st["exp_avg_sq"].numpy(), adam_lr=np.float64(1.6e-4))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
        print(name, {k: float(np.abs(v).max()) for k, v in save.items() if k.startswith("grad")})


if __name__ == "__main__":
    main()
