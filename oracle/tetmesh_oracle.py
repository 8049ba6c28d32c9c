"""CPU oracle for marching tetrahedra -- TEST INFRASTRUCTURE ONLY.

numpy restatement of /root/reference/utils/tetmesh.py:47-138 (_unbatched_marching_tetrahedra, including the
chunked merge of :55-95), pinned by golden vectors produced by importing the reference itself
(tests/golden/make_golden_tetmesh.py -> tests/golden/tetmesh_*.npz)."""
import numpy as np

TRIANGLE_TABLE = np.array([                      # utils/tetmesh.py:23-40
    [-1, -1, -1, -1, -1, -1], [1, 0, 2, -1, -1, -1], [4, 0, 3, -1, -1, -1], [1, 4, 2, 1, 3, 4],
    [3, 1, 5, -1, -1, -1], [2, 3, 0, 2, 5, 3], [1, 4, 0, 1, 5, 4], [4, 2, 5, -1, -1, -1],
    [4, 5, 2, -1, -1, -1], [4, 1, 0, 4, 5, 1], [3, 2, 0, 3, 5, 2], [1, 3, 5, -1, -1, -1],
    [4, 1, 2, 4, 3, 1], [3, 0, 4, -1, -1, -1], [2, 0, 1, -1, -1, -1], [-1, -1, -1, -1, -1, -1]], dtype=np.int64)
NUM_TRIANGLES = np.array([0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0], dtype=np.int64)   # :42
BASE_TET_EDGES = np.array([0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3], dtype=np.int64)              # :43


def _one_chunk(tets, sdf):
    occ_n = sdf > 0
    occ_fx4 = occ_n[tets.reshape(-1)].reshape(-1, 4)
    occ_sum = occ_fx4.sum(-1)
    valid = (occ_sum > 0) & (occ_sum < 4)
    all_edges = tets[valid][:, BASE_TET_EDGES].reshape(-1, 2)
    all_edges = np.sort(all_edges, axis=1)
    if all_edges.shape[0] == 0:
        return np.zeros((0, 2), np.int64), np.zeros((0, 3), np.int64)
    unique_edges, idx_map = np.unique(all_edges, axis=0, return_inverse=True)
    idx_map = idx_map.reshape(-1)
    mask_edges = occ_n[unique_edges.reshape(-1)].reshape(-1, 2).sum(-1) == 1
    mapping = -np.ones(unique_edges.shape[0], np.int64)
    mapping[mask_edges] = np.arange(mask_edges.sum())
    idx_map = mapping[idx_map].reshape(-1, 6)
    interp_v = unique_edges[mask_edges]
    tetindex = (occ_fx4[valid] * (2 ** np.arange(4))[None]).sum(-1)
    ntri = NUM_TRIANGLES[tetindex]
    f1 = np.take_along_axis(idx_map[ntri == 1], TRIANGLE_TABLE[tetindex[ntri == 1]][:, :3], axis=1).reshape(-1, 3)
    f2 = np.take_along_axis(idx_map[ntri == 2], TRIANGLE_TABLE[tetindex[ntri == 2]][:, :6], axis=1).reshape(-1, 3)
    return interp_v, np.concatenate([f1, f2], axis=0)


def marching_tetrahedra(vertices, tets, sdf, scales, chunk_size=32 * 1024 * 1024):
    """Returns ((edge_pos[E,2,3], edge_sdf[E,2,1]), edge_scales[E,2,1], faces[F,3], interp_v[E,2]) like the reference's
    _unbatched_marching_tetrahedra."""
    tets = np.asarray(tets, np.int64)
    sdf = np.asarray(sdf, np.float32)
    if tets.shape[0] > chunk_size:
        n = tets.shape[0] // chunk_size + 1
        size = -(-tets.shape[0] // n)                       # torch.chunk: ceil(len / n) rows per chunk
        merged_ids, merged_faces = None, None
        for c0 in range(0, tets.shape[0], size):
            ids, faces = _one_chunk(tets[c0:c0 + size], sdf)
            if merged_ids is None:
                merged_ids, merged_faces = ids, faces
            else:
                all_edges = np.concatenate([merged_ids, ids], axis=0)
                unique_edges, idx_map = np.unique(all_edges, axis=0, return_inverse=True)
                idx_map = idx_map.reshape(-1)
                f0 = idx_map[merged_faces.reshape(-1)].reshape(-1, 3)
                f1 = idx_map[faces.reshape(-1) + merged_ids.shape[0]].reshape(-1, 3)
                merged_ids, merged_faces = unique_edges, np.concatenate([f0, f1], axis=0)
        interp_v, faces = merged_ids, merged_faces
    else:
        interp_v, faces = _one_chunk(tets, sdf)
    v = np.asarray(vertices, np.float32)
    sc = np.asarray(scales, np.float32).reshape(-1, 1)
    flat = interp_v.reshape(-1)
    return (v[flat].reshape(-1, 2, 3), sdf[flat].reshape(-1, 2, 1)), sc[flat].reshape(-1, 2, 1), faces, interp_v
