import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _util, gof_synth
from diff_gaussian_rasterization import _C as ours
dev = torch.device("cuda")
cam, gs = gof_synth.make_scene(sys.argv[1] if len(sys.argv) > 1 else "C3", view=1)
fa = _util.fwd_args(cam, gs, dev)
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    o = ours.rasterize_gaussians(*fa)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"iter {i}: call returned after {1e3*(t1-t0):.3f} ms, gpu done after {1e3*(t2-t0):.3f} ms", file=sys.stderr, flush=True)
