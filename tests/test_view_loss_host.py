"""CPU: the per-view loss kernel source (csrc/view_loss.cuh), compiled for the host and run phase by phase
(tests/hostmath/view_loss_host.cpp), against the reference-generated golden vectors and the oracle."""
import ctypes
import glob
import math
import os
import subprocess

import numpy as np
import pytest

import loss_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = sorted(glob.glob(os.path.join(HERE, "golden", "loss_*.npz")))


@pytest.fixture(scope="module")
def hm():
    d = os.path.join(HERE, "hostmath")
    lib, src = os.path.join(d, "libviewloss_host.so"), os.path.join(d, "view_loss_host.cpp")
    hdr = os.path.join(HERE, "..", "gaussian-opacity-fields_b200", "csrc", "view_loss.cuh")
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-x", "c++", src, "-o", lib])
    return ctypes.CDLL(lib)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def run_host(hm, render, gt, wvt, tanfovx, tanfovy, lambdas, need_grad=True):
    _, H, W = render.shape
    c2w = np.linalg.inv(np.asarray(wvt, np.float64).T)
    R9 = np.ascontiguousarray(c2w[:3, :3], np.float32).reshape(9)
    g = np.array([math.exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)], np.float32)
    g = (g / g.sum()).astype(np.float32)
    terms = np.zeros(5, np.float32)
    grad = np.zeros_like(render) if need_grad else None
    render, gt = np.ascontiguousarray(render, np.float32), np.ascontiguousarray(gt, np.float32)
    hm.hm_view_loss(W, H, _p(render), _p(gt), _p(R9), ctypes.c_float(W / (2 * tanfovx)), ctypes.c_float(H / (2 * tanfovy)), _p(g),
                    ctypes.c_float(lambdas[0]), ctypes.c_float(lambdas[1]), ctypes.c_float(lambdas[2]), _p(terms),
                    _p(grad) if need_grad else None)
    return terms, grad


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[:-4] for p in FIX])
def test_kernel_source_matches_reference_goldens(hm, path):
    fx = np.load(path)
    terms, grad = run_host(hm, fx["render"], fx["gt"], fx["world_view_transform"], float(fx["tanfovx"]), float(fx["tanfovy"]), fx["lambdas"])
    for i, k in enumerate(("Ll1", "ssim", "depth_normal_loss", "distortion_loss", "loss")):
        assert abs(float(terms[i]) - float(fx[k])) <= 1e-5 * max(1.0, abs(float(fx[k]))), k
    r = fx["grad"].astype(np.float64)
    for ch in range(9):
        den = max(np.abs(r[ch]).max(), 1e-12)
        assert np.abs(grad[ch] - r[ch]).max() / den < 1e-4, f"channel {ch}"


def test_kernel_source_vs_oracle_on_ragged_image(hm):
    """Sizes that are not multiples of the tile, smaller than the window in one dimension, several tiles in the other."""
    rng = np.random.default_rng(7)
    import gof_synth
    for (W, H) in ((37, 9), (16, 16), (50, 35)):
        cam = gof_synth.make_camera(W, H, view=12)
        render = rng.uniform(0, 1, size=(9, H, W)).astype(np.float32)
        render[3:6] -= 0.5
        render[6] = 2.0 + render[6]
        gt = rng.uniform(0, 1, size=(3, H, W)).astype(np.float32)
        lambdas = (0.2, 0.05, 100.0)
        terms, grad = run_host(hm, render, gt, cam.world_view_transform.numpy(), cam.tanfovx, cam.tanfovy, lambdas)
        out = loss_oracle.view_loss(render, gt, cam.world_view_transform.numpy(), cam.tanfovx, cam.tanfovy, lambdas)
        assert abs(float(terms[4]) - out["loss"]) <= 1e-5 * max(1.0, abs(out["loss"]))
        for ch in range(9):
            den = max(np.abs(out["grad"][ch]).max(), 1e-12)
            assert np.abs(grad[ch] - out["grad"][ch]).max() / den < (2e-3 if ch == 6 else 1e-4), f"{W}x{H} channel {ch}"
        # values only (no gradient buffer)
        t2, _ = run_host(hm, render, gt, cam.world_view_transform.numpy(), cam.tanfovx, cam.tanfovy, lambdas, need_grad=False)
        assert np.array_equal(t2, terms)
