"""CPU: the parameter prologue / epilogue source (csrc/param_ops.cuh) compiled for the host, against golden vectors
generated from the reference's own Python (tests/golden/make_golden_params.py): activations with the 3D filter
(scene/gaussian_model.py:152-194), their backward, and torch.optim.Adam(eps=1e-15) (:360)."""
import ctypes
import glob
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = sorted(glob.glob(os.path.join(HERE, "golden", "params_*.npz")))


@pytest.fixture(scope="module")
def hm():
    d = os.path.join(HERE, "hostmath")
    lib, src = os.path.join(d, "libparamops_host.so"), os.path.join(d, "param_ops_host.cpp")
    hdr = os.path.join(HERE, "..", "gaussian-opacity-fields_b200", "csrc", "param_ops.cuh")
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-x", "c++", src, "-o", lib])
    return ctypes.CDLL(lib)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[:-4] for p in FIX])
def test_activation_and_backward(hm, path):
    fx = np.load(path)
    P = fx["raw_scaling"].shape[0]
    s, q, o, f = (np.ascontiguousarray(fx[k], np.float32) for k in ("raw_scaling", "raw_rotation", "raw_opacity", "filter_3D"))
    scales, rot, op = np.zeros((P, 3), np.float32), np.zeros((P, 4), np.float32), np.zeros((P, 1), np.float32)
    hm.hm_activate(P, _p(s), _p(q), _p(o), _p(f), _p(scales), _p(rot), _p(op))
    assert _rel(scales, fx["out_scales"]) < 2e-6 and _rel(rot, fx["out_rotations"]) < 2e-6 and _rel(op, fx["out_opacities"]) < 5e-6
    gs, gr, go = (np.ascontiguousarray(fx[k], np.float32) for k in ("up_scales", "up_rotations", "up_opacities"))
    ds, dq, do = np.zeros((P, 3), np.float32), np.zeros((P, 4), np.float32), np.zeros((P, 1), np.float32)
    hm.hm_activate_backward(P, _p(s), _p(q), _p(o), _p(f), _p(gs), _p(gr), _p(go), _p(ds), _p(dq), _p(do))
    assert _rel(ds, fx["grad_scaling"]) < 2e-5 and _rel(do, fx["grad_opacity"]) < 2e-5
    ok = np.linalg.norm(q, axis=1) > 1e-6                       # the degenerate quaternion separately: gradient = g / eps
    assert _rel(dq[ok], fx["grad_rotation"][ok]) < 2e-5
    assert _rel(dq[~ok], fx["grad_rotation"][~ok]) < 1e-5
    # features: shs = cat(f_dc, f_rest) -> the gradient is split back
    up = fx["up_shs"]
    assert np.array_equal(up[:, :1], fx["grad_features_dc"]) and np.array_equal(up[:, 1:], fx["grad_features_rest"])


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[:-4] for p in FIX])
def test_adam_three_steps(hm, path):
    fx = np.load(path)
    p = np.ascontiguousarray(fx["adam_p0"], np.float32).copy()
    m, v = np.zeros_like(p), np.zeros_like(p)
    for t, g in enumerate(fx["adam_grads"], start=1):
        g = np.ascontiguousarray(g, np.float32)
        hm.hm_adam(ctypes.c_long(p.size), _p(p), _p(m), _p(v), _p(g), ctypes.c_double(float(fx["adam_lr"])), ctypes.c_double(0.9),
                   ctypes.c_double(0.999), ctypes.c_double(1e-15), t)
    assert _rel(m, fx["adam_m"]) < 1e-6 and _rel(v, fx["adam_v"]) < 1e-6
    assert float(np.abs(p - fx["adam_p"]).max()) < 2e-6 * float(fx["adam_lr"]) / 1.6e-4 + 1e-7
