#!/usr/bin/env python
"""Golden vectors for the per-view training loss (train.py:151-188 of the reference): generated HERE by executing the
reference's own Python (utils/loss_utils.py, utils/depth_utils.py) on the CPU.  The two files are read from
/root/reference and exec'ed IN MEMORY with `.cuda()` / `device='cuda'` neutralised (this container has no GPU); nothing of
them is written to the repo.  Output: tests/golden/loss_*.npz = inputs, every loss term, the total, and the gradient of
the total with respect to the 9-channel render (autograd, float32 like the reference)."""
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "gaussian-opacity-fields_b200"))
import gof_synth  # noqa: E402


def load_ref_module(rel):
    src = open(os.path.join(REF, rel)).read()
    src = src.replace(".cuda()", "").replace("device='cuda'", "device='cpu'")
    m = types.ModuleType("ref_" + os.path.basename(rel)[:-3])
    exec(compile(src, rel, "exec"), m.__dict__)
    return m


class View:   # the attributes depth_to_normal reads (scene/cameras.py:28-29,40-41,56)
    def __init__(self, cam):
        self.world_view_transform = cam.world_view_transform
        self.image_width, self.image_height = cam.image_width, cam.image_height
        self.FoVx = 2.0 * math.atan(cam.tanfovx)
        self.FoVy = 2.0 * math.atan(cam.tanfovy)


CASES = {
    "loss_a": dict(W=72, H=48, seed=1, view=5, lambdas=(0.2, 0.05, 100.0)),      # arguments/__init__.py:93-95
    "loss_b": dict(W=40, H=64, seed=2, view=23, lambdas=(0.2, 0.0, 0.0)),        # before iteration 15000 (:96-97)
    "loss_c": dict(W=33, H=21, seed=3, view=40, lambdas=(0.35, 0.3, 10.0)),      # odd sizes smaller than two windows
}


def make_inputs(cfg):
    g = torch.Generator().manual_seed(cfg["seed"])
    H, W = cfg["H"], cfg["W"]
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    render = torch.zeros(9, H, W)
    gt = (0.5 + 0.4 * torch.sin(6.0 * xx + 3.0 * yy)[None] * torch.tensor([1.0, 0.7, -0.8])[:, None, None]
          + 0.05 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    render[:3] = (gt + 0.08 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    render[3:6] = torch.randn(3, H, W, generator=g) * 0.6                      # un-normalised alpha-weighted normals
    render[3:6, :2, :3] = 0.0                                                  # empty pixels: zero normal
    render[6] = 3.0 + 0.5 * torch.sin(4.0 * xx) * torch.cos(5.0 * yy) + 0.02 * torch.randn(H, W, generator=g)   # median depth
    render[6, -3:, -4:] = 0.0                                                  # holes
    render[7] = torch.rand(H, W, generator=g)
    render[8] = torch.rand(H, W, generator=g) * 1e-2
    return render.contiguous(), gt.contiguous()


def main():
    lu, du = load_ref_module("utils/loss_utils.py"), load_ref_module("utils/depth_utils.py")
    for name, cfg in CASES.items():
        cam = gof_synth.make_camera(cfg["W"], cfg["H"], view=cfg["view"])
        view = View(cam)
        render0, gt = make_inputs(cfg)
        lam_dssim, lam_dn, lam_dist = cfg["lambdas"]
        rendering = render0.clone().requires_grad_(True)
        # ---- train.py:151-188, line by line ----
        image = rendering[:3, :, :]
        Ll1 = lu.l1_loss(image, gt)
        ssim_v = lu.ssim(image, gt)
        rgb_loss = (1.0 - lam_dssim) * Ll1 + lam_dssim * (1.0 - ssim_v)
        distortion_loss = rendering[8, :, :].mean()
        depth = rendering[6, :, :]
        depth_normal, _ = du.depth_to_normal(view, depth[None, ...])
        depth_normal = depth_normal.permute(2, 0, 1)
        render_normal = torch.nn.functional.normalize(rendering[3:6, :, :], p=2, dim=0)
        c2w = (view.world_view_transform.T).inverse()
        normal2 = c2w[:3, :3] @ render_normal.reshape(3, -1)
        render_normal_world = normal2.reshape(3, *render_normal.shape[1:])
        normal_error = 1 - (render_normal_world * depth_normal).sum(dim=0)
        depth_normal_loss = normal_error.mean()
        loss = rgb_loss + depth_normal_loss * lam_dn + distortion_loss * lam_dist
        loss.backward()
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), render=render0.numpy(), gt=gt.numpy(),
            world_view_transform=cam.world_view_transform.numpy(), tanfovx=np.float64(cam.tanfovx), tanfovy=np.float64(cam.tanfovy),
            lambdas=np.array(cfg["lambdas"], np.float64), Ll1=Ll1.detach().numpy(), ssim=ssim_v.detach().numpy(),
            distortion_loss=distortion_loss.detach().numpy(), depth_normal_loss=depth_normal_loss.detach().numpy(),
            depth_normal=depth_normal.detach().numpy(), loss=loss.detach().numpy(), grad=rendering.grad.numpy())
        print(name, float(loss), float(Ll1), float(ssim_v), float(depth_normal_loss), float(distortion_loss))


if __name__ == "__main__":
    main()
