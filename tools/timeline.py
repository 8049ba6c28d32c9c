#!/usr/bin/env python
"""GPU timeline of the bench's device loop from the library's event brackets: python tools/timeline.py [C3] [steps]
Prints every launch with the idle gap in front of it, plus host wall time per phase."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _util, gof_synth, gof_dp
from diff_gaussian_rasterization import _C
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda")
scenes = [gof_synth.make_scene(name, view=v) for v in range(4)]
gs = scenes[0][1]
fas = [_util.fwd_args(c, gs, dev) for c, _ in scenes]
H, W = scenes[0][0].image_height, scenes[0][0].image_width
P = gs["means3D"].shape[0]
grad = torch.randn(9, H, W, device=dev)
bucket = gof_dp.GradBucket(P, 16, dev)

def step(i, log=None):
    fa = fas[i % 4]
    t0 = time.perf_counter()
    R, color, radii, geom, binning, img = _C.rasterize_gaussians(*fa)
    t1 = time.perf_counter()
    # (the backward writes every element of its outputs: no zero fill)
    g = _C.rasterize_gaussians_backward(*_util.bwd_args(fa, radii, geom, R, binning, img, grad), _out=bucket.views)
    t2 = time.perf_counter()
    if log is not None:
        log.append((1e3 * (t1 - t0), 1e3 * (t2 - t1)))

for i in range(4):
    step(i)
torch.cuda.synchronize()
# untimed-bracket run for the true step time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
log = []
e0.record()
for i in range(10):
    step(i, log)
e1.record(); torch.cuda.synchronize()
print(f"plain loop: {e0.elapsed_time(e1) / 10:.3f} ms/step; host fwd/bwd call ms:", " ".join(f"{a:.2f}/{b:.2f}" for a, b in log))
_C.profile_reset(); _C.profile_enable(True)
e0.record()
for i in range(steps):
    step(i)
e1.record(); torch.cuda.synchronize()
_C.profile_enable(False)
print(f"bracketed loop: {e0.elapsed_time(e1) / steps:.3f} ms/step")
tl = _C.profile_timeline()
prev_end, busy = None, 0.0
for n, a, b in tl:
    gap = 0.0 if prev_end is None else a - prev_end
    busy += b - a
    flag = "  <-- idle" if gap > 0.03 else ""
    print(f"{n:20s} start {a:9.3f}  dur {b - a:7.3f}  gap {gap:7.3f}{flag}")
    prev_end = b
print(f"busy {busy:.3f} ms of {tl[-1][2] - tl[0][1]:.3f} ms")
