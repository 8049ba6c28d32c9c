"""CPU: the culling box (gof_cull_bbox in csrc/gof_math.cuh, host twin) is conservative: every pixel at which the
reference's pair test accepts a Gaussian (oracle_alpha_map) lies inside that Gaussian's box -- including strongly
anisotropic Gaussians, large and tiny footprints, and a camera inside the cloud."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

import gof_oracle
import gof_synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hm():
    d = os.path.join(HERE, "hostmath")
    lib, src = os.path.join(d, "libhostmath.so"), os.path.join(d, "hostmath.cpp")
    hdr = os.path.join(HERE, "..", "gaussian-opacity-fields_b200", "csrc", "gof_math.cuh")
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-x", "c++", src, "-o", lib])
    return ctypes.CDLL(lib)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


CASES = [
    dict(P=500, W=96, H=64, seed=1, sigma_px=2.0, radius=4.0, aniso=(0.3, 1.0)),
    dict(P=300, W=80, H=60, seed=2, sigma_px=0.6, radius=4.0, aniso=(0.3, 1.0)),     # sub-pixel Gaussians, C ~ 1e6
    dict(P=200, W=64, H=48, seed=3, sigma_px=12.0, radius=4.0, aniso=(0.3, 1.0)),    # screen-filling
    dict(P=400, W=96, H=64, seed=4, sigma_px=3.0, radius=4.0, aniso=(0.01, 1.0)),    # needles / pancakes 100:1
    dict(P=400, W=96, H=64, seed=5, sigma_px=3.0, radius=1.0, aniso=(0.1, 1.0)),     # camera inside the cloud
]


@pytest.mark.parametrize("case", CASES, ids=[f"case{i}" for i in range(len(CASES))])
def test_box_contains_every_accepted_pixel(hm, case):
    cam = gof_synth.make_camera(case["W"], case["H"], view=case["seed"] * 7, radius=case["radius"])
    gs = gof_synth.make_gaussians(case["P"], case["seed"], cam.focal_x, sigma_px=case["sigma_px"])
    lo, hi = case["aniso"]
    g = torch.Generator().manual_seed(case["seed"] + 99)
    an = torch.exp(torch.rand(case["P"], 3, generator=g) * np.log(hi / lo) + np.log(lo))
    gs["scales"] = (gs["scales"].max(dim=1, keepdim=True).values * an).contiguous()
    sc = gof_oracle.scene_from_synth(cam, gs)
    st = gof_oracle.preprocess(sc)
    vis = np.nonzero(st["radii"] > 0)[0]
    assert len(vis) > 50
    W, H = case["W"], case["H"]
    tight = nonempty = 0
    box = np.zeros(4, np.int32)
    for gid in vis:
        v = np.ascontiguousarray(st["view2gaussian"][gid])
        op = float(st["conic_opacity"][gid, 3])
        scale = np.ascontiguousarray(sc.arr["scales"][gid])
        hm.hm_bbox(_p(v), ctypes.c_float(op), _p(scale), W, H, ctypes.c_float(sc.tan_fovx), ctypes.c_float(sc.tan_fovy), _p(box))
        amap = gof_oracle.alpha_map(W, H, sc.tan_fovx, sc.tan_fovy, v, op)
        ys, xs = np.nonzero(amap > 0)
        if len(xs) == 0:
            continue
        nonempty += 1
        assert box[0] <= xs.min() and xs.max() <= box[2] and box[1] <= ys.min() and ys.max() <= box[3], \
            f"gaussian {gid}: accepted pixels x[{xs.min()},{xs.max()}] y[{ys.min()},{ys.max()}] outside box {box.tolist()}"
        if box[0] > -30000:
            tight += 1
    assert nonempty > 20
    if case["radius"] >= 4.0:
        assert tight > 0.5 * nonempty, "the box should be finite for most Gaussians in front of the camera"
