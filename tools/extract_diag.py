#!/usr/bin/env python
"""Where does the host time of integrate / marching tets go? wall vs event time per call, allocator statistics."""
import os, sys, time, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _util, gof_synth, gof_tetmesh
from diff_gaussian_rasterization import _C as ours
dev = torch.device("cuda")
cam, gs = gof_synth.make_scene("C3", view=3)
P = gs["means3D"].shape[0]
fa = _util.fwd_args(cam, gs, dev)
for n_pts in (2_000_000, 9_000_000):
    g = torch.Generator().manual_seed(4)
    idx = torch.randint(0, P, (n_pts,), generator=g)
    pts = (gs["means3D"][idx] + gs["scales"][idx] * 3.0 * (torch.rand(n_pts, 3, generator=g) * 2 - 1)).contiguous().to(dev)
    ia = (fa[0], pts) + tuple(fa[1:])
    for it in range(6):
        torch.cuda.synchronize()
        st0 = torch.cuda.memory_stats()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        out = ours.integrate_gaussians_to_points(*ia)
        b.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        st1 = torch.cuda.memory_stats()
        print(f"integrate {n_pts} it{it}: host {1e3*(t1-t0):.2f} ms, done {1e3*(t2-t0):.2f} ms, events {a.elapsed_time(b):.2f} ms, "
              f"cudaMalloc +{st1['num_device_alloc']-st0['num_device_alloc']} cudaFree +{st1['num_device_free']-st0['num_device_free']} "
              f"reserved {st1['reserved_bytes.all.current']/2**20:.0f} MiB", flush=True)
        del out
rng = np.random.default_rng(5)
V, T = 400_000, 2_600_000
v = rng.uniform(-1, 1, size=(V, 3)).astype(np.float32)
a_ = rng.integers(0, V, size=T)
tets = np.stack([a_, (a_ + rng.integers(1, 50, size=T)) % V, (a_ + rng.integers(50, 400, size=T)) % V, (a_ + rng.integers(400, 3000, size=T)) % V], axis=1).astype(np.int64)
sdf = (0.8 - np.linalg.norm(v, axis=1) + 0.05 * rng.standard_normal(V)).astype(np.float32)
sc = rng.uniform(0.01, 0.1, size=(V, 1)).astype(np.float32)
t = lambda x: torch.from_numpy(x).to(dev)
tv, tt, ts_, tsc = t(v), t(tets), t(sdf), t(sc)
for it in range(6):
    torch.cuda.synchronize(); st0 = torch.cuda.memory_stats()
    ours.profile_reset(); ours.profile_enable(it == 5)
    t0 = time.perf_counter()
    out = gof_tetmesh._unbatched_marching_tetrahedra(tv, tt, ts_, tsc)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    st1 = torch.cuda.memory_stats()
    print(f"tetmesh it{it}: host {1e3*(t1-t0):.2f} ms, done {1e3*(t2-t0):.2f} ms, cudaMalloc +{st1['num_device_alloc']-st0['num_device_alloc']} "
          f"cudaFree +{st1['num_device_free']-st0['num_device_free']}", flush=True)
ours.profile_enable(False)
print({k: round(v[1], 3) for k, v in ours.profile_report().items()})
