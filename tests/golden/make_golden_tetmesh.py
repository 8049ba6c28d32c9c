#!/usr/bin/env python
"""Golden vectors for marching tetrahedra, produced by IMPORTING the reference's own implementation
(/root/reference/utils/tetmesh.py, pure torch, runs on CPU in the build container) on seeded inputs.
Writes tests/golden/tetmesh_*.npz (inputs + the reference's outputs).  /root/reference is not available on the
GPU box; only the committed vectors travel."""
import importlib.util
import os
import sys

import numpy as np
import torch
from scipy.spatial import Delaunay

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_tetmesh", "/root/reference/utils/tetmesh.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

CASES = {"tetmesh_small": dict(n=400, seed=1, kind="sphere"), "tetmesh_noisy": dict(n=1500, seed=2, kind="noisy"),
         "tetmesh_allout": dict(n=200, seed=3, kind="allout")}
for name, c in CASES.items():
    rng = np.random.default_rng(c["seed"])
    v = rng.uniform(-1, 1, size=(c["n"], 3)).astype(np.float32)
    tets = Delaunay(v).simplices.astype(np.int64)
    if c["kind"] == "sphere":
        sdf = (0.7 - np.linalg.norm(v, axis=1)).astype(np.float32)
    elif c["kind"] == "noisy":
        sdf = (0.6 - np.linalg.norm(v * [1.0, 0.7, 1.3], axis=1) + 0.15 * rng.standard_normal(c["n"])).astype(np.float32)
    else:
        sdf = -np.ones(c["n"], np.float32)
    scales = rng.uniform(0.01, 0.1, size=(c["n"], 1)).astype(np.float32)
    verts, vscales, faces, interp_v = ref.marching_tetrahedra(torch.from_numpy(v)[None], torch.from_numpy(tets), torch.from_numpy(sdf)[None],
                                                              torch.from_numpy(scales)[None])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), vertices=v, tets=tets, sdf=sdf, scales=scales,
                        edge_pos=verts[0][0].numpy(), edge_sdf=verts[0][1].numpy(), edge_scales=vscales[0].numpy(),
                        faces=faces[0].numpy(), interp_v=interp_v[0].numpy())
    print(name, "tets", tets.shape[0], "edges", interp_v[0].shape[0], "faces", faces[0].shape[0])

# chunked merge path (utils/tetmesh.py:55-95): the reference hard-codes 32 Mi tets per chunk; to exercise that code on a
# fixture-sized input its source is executed here with only that constant lowered to 1000 (nothing else changed).
src = open("/root/reference/utils/tetmesh.py").read().replace("chunk_size = 32 * 1024 * 1024", "chunk_size = 1000")
ns = {}
exec(compile(src, "ref_tetmesh_chunked", "exec"), ns)
rng = np.random.default_rng(7)
v = rng.uniform(-1, 1, size=(900, 3)).astype(np.float32)
tets = Delaunay(v).simplices.astype(np.int64)
sdf = (0.65 - np.linalg.norm(v, axis=1) + 0.1 * rng.standard_normal(900)).astype(np.float32)
scales = rng.uniform(0.01, 0.1, size=(900, 1)).astype(np.float32)
verts, vscales, faces, interp_v = ns["marching_tetrahedra"](torch.from_numpy(v)[None], torch.from_numpy(tets), torch.from_numpy(sdf)[None],
                                                            torch.from_numpy(scales)[None])
np.savez_compressed(os.path.join(HERE, "tetmesh_chunked1000.npz"), vertices=v, tets=tets, sdf=sdf, scales=scales,
                    edge_pos=verts[0][0].numpy(), edge_sdf=verts[0][1].numpy(), edge_scales=vscales[0].numpy(),
                    faces=faces[0].numpy(), interp_v=interp_v[0].numpy(), chunk_size=np.int64(1000))
print("tetmesh_chunked1000 tets", tets.shape[0], "edges", interp_v[0].shape[0], "faces", faces[0].shape[0])
