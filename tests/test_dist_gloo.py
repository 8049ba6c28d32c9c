"""CPU, world_size 2 over gloo: the view-parallel plumbing (gof_dp) -- flat gradient bucket all-reduce equals the
sum of per-rank gradients, densification statistics reduce with SUM/SUM/SUM/MAX, view schedule covers the ring."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "..", "gaussian-opacity-fields_b200"))
    import gof_dp
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, M = 257, 16
    b = gof_dp.GradBucket(P, M, "cpu")
    assert sum(v.numel() for v in b.views.values()) == P * (59 + 5)      # gradients + (dens_sum 3 | dens_max 2)
    g = torch.Generator().manual_seed(rank)
    local = {}
    for name, v in b.views.items():
        assert v.is_contiguous()
        v.copy_(torch.randn(v.shape, generator=g))
        local[name] = v.clone()
    b.all_reduce()
    dm2 = torch.randn(P, 3, generator=g)
    radii = torch.randint(0, 5, (P,), generator=g, dtype=torch.int32)
    st = gof_dp.all_reduce_densification_stats(gof_dp.densification_stats(dm2, radii))
    # plain numpy in the queue: tensors would be shared by fd and the worker exits before the parent reads
    q.put((rank, {k: v.numpy().copy() for k, v in local.items()}, {k: v.numpy().copy() for k, v in b.views.items()},
           dm2.numpy().copy(), radii.numpy().copy(), st.numpy().copy(), [gof_dp.view_for(s, rank, world) for s in range(40)]))
    dist.destroy_process_group()


def test_bucket_and_stats_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = [(r[0], {k: torch.from_numpy(v) for k, v in r[1].items()}, {k: torch.from_numpy(v) for k, v in r[2].items()},
            torch.from_numpy(r[3]), torch.from_numpy(r[4]), torch.from_numpy(r[5]), r[6]) for r in res]
    for name in res[0][1]:
        total = torch.maximum(res[0][1][name], res[1][1][name]) if name == "dens_max" else res[0][1][name] + res[1][1][name]
        for r in range(world):
            assert torch.equal(res[r][2][name], total), name     # dens_max rides in the MAX tail of the same bucket
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gaussian-opacity-fields_b200"))
    import gof_dp
    s0, s1 = gof_dp.densification_stats(res[0][3], res[0][4]), gof_dp.densification_stats(res[1][3], res[1][4])
    want = torch.cat([s0[:, :3] + s1[:, :3], torch.maximum(s0[:, 3], s1[:, 3])[:, None]], dim=1)
    for r in range(world):
        assert torch.allclose(res[r][5], want)
    views = sorted(res[0][6] + res[1][6])
    assert views[:64] == sorted(list(range(64)))[:64] or set(views) == set(range(64))


def test_bucket_single_process_noop():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gaussian-opacity-fields_b200"))
    import gof_dp
    b = gof_dp.GradBucket(10, 16, "cpu")
    b.views["dsh"].fill_(1.0)
    assert b.all_reduce() is None
    assert float(b.flat.sum()) == 10 * 48
    assert gof_dp.GradBucket(10, 16, "cpu", with_stats=False).n_sum == gof_dp.GradBucket(10, 16, "cpu", with_stats=False).numel
    b.zero_()
    assert float(b.flat.abs().sum()) == 0.0


def test_bucket_layout_and_padding():
    """Views appear in the documented order, every one starts on a 256-byte boundary for ANY P (k_preprocess_backward stores
    dL_drot / dL_dsh with 128-bit stores), the padding is never written, and the peer exchange is a no-op without a group."""
    import gof_dp
    for P in (1, 3, 10, 1001, 100_003):
        b = gof_dp.GradBucket(P, 16, "cpu")
        assert b.flat.numel() % 64 == 0 and b.flat.numel() == b.numel
        last = -1
        assert b.n_sum == (b.views["dens_max"].data_ptr() - b.flat.data_ptr()) // 4      # SUM region ends where the MAX tail starts
        for name, per in (("dmeans3D", 3), ("dsh", 48), ("dopacity", 1), ("dscales", 3), ("drot", 4), ("dens_sum", 3), ("dens_max", 2)):
            v = b.views[name]
            off = (v.data_ptr() - b.flat.data_ptr()) // 4
            assert v.numel() == per * P and v.is_contiguous()
            assert off % 64 == 0 and off > last, name
            last = off + v.numel() - 1
        assert last < b.flat.numel()
        b.views["dsh"].fill_(2.0)
        assert float(b.flat.sum()) == 2.0 * 48 * P
        assert b.enable_peer_exchange() is b and b.exchange == "nccl"
        assert b.all_reduce() is None
        b.close()                                   # no-op outside peer mode
