#!/usr/bin/env python
"""CPU experiment: how many (8x4 warp rectangle, Gaussian) visits does a given cull box produce, against the visits in
which at least one pixel really passes the alpha test?  Scaled-down C3 (same footprint in pixels)."""
import ctypes, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussian-opacity-fields_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gof_oracle, gof_synth
d = os.path.join(ROOT, "tests", "hostmath")
lib, src = os.path.join(d, "libhostmath.so"), os.path.join(d, "hostmath.cpp")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-x", "c++", src, "-o", lib])
hm = ctypes.CDLL(lib)
_p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
W, H, P = 480, 272, int(sys.argv[1]) if len(sys.argv) > 1 else 4000
cam = gof_synth.make_camera(W, H, view=1)
gs = gof_synth.make_gaussians(P, 2, cam.focal_x)
sc = gof_oracle.scene_from_synth(cam, gs)
st = gof_oracle.preprocess(sc)
vis = np.nonzero(st["radii"] > 0)[0]
box = np.zeros(4, np.int32)
def rects(x0, y0, x1, y1):
    x0, y0, x1, y1 = max(x0, 0), max(y0, 0), min(x1, W - 1), min(y1, H - 1)
    if x0 > x1 or y0 > y1: return 0
    return (x1 // 8 - x0 // 8 + 1) * (y1 // 4 - y0 // 4 + 1)
n_box = n_true = n_tightbb = n_rect3s = 0
for gid in vis:
    v = np.ascontiguousarray(st["view2gaussian"][gid]); op = float(st["conic_opacity"][gid, 3])
    scale = np.ascontiguousarray(sc.arr["scales"][gid])
    hm.hm_bbox(_p(v), ctypes.c_float(op), _p(scale), W, H, ctypes.c_float(sc.tan_fovx), ctypes.c_float(sc.tan_fovy), _p(box))
    n_box += rects(*box)
    amap = gof_oracle.alpha_map(W, H, sc.tan_fovx, sc.tan_fovy, v, op)
    ys, xs = np.nonzero(amap > 0)
    if len(xs):
        n_true += len(set(zip((xs // 8).tolist(), (ys // 4).tolist())))
        n_tightbb += rects(xs.min(), ys.min(), xs.max(), ys.max())
print(f"visible {len(vis)}: visits with the shipped box {n_box}, with the exact pixel AABB {n_tightbb}, with >=1 passing pixel {n_true}")
