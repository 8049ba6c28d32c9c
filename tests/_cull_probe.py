"""Helper for test_gpu_culling.py: renders the scenes stored in an .npz (written ONCE by the parent test, so that both
GOF_CULL modes see bit-identical inputs -- torch's vectorised CPU exp/sigmoid may differ by an ulp between processes)
and dumps the outputs.  Usage: _cull_probe.py <inputs.npz> <outputs.npz>, GOF_CULL=0/1 in the environment."""
import sys
import numpy as np
import torch
import _util
import gof_synth
from diff_gaussian_rasterization import _C

CASES = [("a", dict(P=40000, width=640, height=360, seed=3), 2, 2.0, (0.3, 1.0), 4.0),
         ("b", dict(P=20000, width=320, height=200, seed=4), 9, 6.0, (0.02, 1.0), 2.0),   # needles / pancakes
         ("c", dict(P=20000, width=320, height=200, seed=5), 5, 0.7, (0.3, 1.0), 4.0)]    # sub-pixel
KEYS = ("means3D", "scales", "rotations", "opacities", "shs")


def make_inputs(path):
    out = {}
    for tag, cfg, view, sigma, aniso, radius in CASES:
        cam = gof_synth.make_camera(cfg["width"], cfg["height"], view=view, radius=radius)
        gs = gof_synth.make_gaussians(cfg["P"], cfg["seed"], cam.focal_x, sigma_px=sigma)
        g = torch.Generator().manual_seed(cfg["seed"] + 50)
        an = torch.exp(torch.rand(cfg["P"], 3, generator=g) * np.log(aniso[1] / aniso[0]) + np.log(aniso[0]))
        gs["scales"] = (gs["scales"].max(dim=1, keepdim=True).values * an).contiguous()
        for k in KEYS:
            out[f"{tag}_{k}"] = gs[k].numpy()
    np.savez(path, **out)


def main(in_path, out_path):
    dev = torch.device("cuda")
    inp = np.load(in_path)
    res = {}
    for tag, cfg, view, sigma, aniso, radius in CASES:
        cam = gof_synth.make_camera(cfg["width"], cfg["height"], view=view, radius=radius)
        gs = {k: torch.from_numpy(inp[f"{tag}_{k}"]) for k in KEYS}
        gs["sh_degree"] = 3
        fa = _util.fwd_args(cam, gs, dev)
        R, color, radii, geom, binning, img = _C.rasterize_gaussians(*fa)
        st = _C.export_state(cfg["P"], cfg["width"], cfg["height"], R, geom, binning, img, radii)
        grad = torch.randn(9, cfg["height"], cfg["width"], generator=torch.Generator().manual_seed(1)).to(dev)
        grads = _C.rasterize_gaussians_backward(*_util.bwd_args(fa, radii, geom, R, binning, img, grad))
        res[tag + "_color"] = color.cpu().numpy()
        res[tag + "_ncontrib"] = st["n_contrib"].cpu().numpy()
        res[tag + "_accum"] = st["accum_alpha"].cpu().numpy()
        for n, t in zip(["dmeans2D", "dcolors", "dopacity", "dmeans3D", "dcov3D", "dsh", "dscales", "drot", "dv2g"], grads):
            res[tag + "_" + n] = t.cpu().numpy()
    np.savez(out_path, **res)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
