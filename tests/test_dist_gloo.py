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


def _worker_factored(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "..", "gaussian-opacity-fields_b200"))
    import gof_dp
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, M = 203, 16
    b = gof_dp.GradBucket(P, M, "cpu", factor_sh=True)
    assert "dsh_rgb" in b.views and b.views["dsh"].shape == (P, M, 3) and b.n_reduce == b.n_sum + 2 * P // 64 * 64 + 64 * (2 * P % 64 != 0)
    plane = (P + 63) // 64 * 64
    assert b.numel == b.n_reduce + world * (64 + 3 * plane) and b.views["dsh_rgb"].shape == (3, plane)
    g = torch.Generator().manual_seed(10 + rank)
    means = torch.randn(P, 3, generator=torch.Generator().manual_seed(5)) * 2     # the same Gaussians on every rank
    cam = torch.tensor([3.0 + rank, -1.0, 0.5 * rank])
    degree = 3 - rank                                                             # (ranks may be at different degrees only in a test)
    local = {}
    for name in ("dmeans3D", "dopacity", "dscales", "drot", "dens_sum", "dens_max", "dsh_rgb"):
        b.views[name].copy_(torch.randn(b.views[name].shape, generator=g))
        local[name] = b.views[name].clone()
    b.views["dsh_rgb"][:, ::3] = 0.0
    b.views["dsh_rgb"][:, P:] = 0.0
    local["dsh_rgb"] = b.views["dsh_rgb"][:, :P].t().contiguous()        # (P,3)
    b.views["sh_hdr"][:4] = torch.tensor([cam[0], cam[1], cam[2], float(degree)])
    b.all_reduce(means3D=means)
    q.put((rank, {k: v.numpy().copy() for k, v in local.items()}, {k: v.numpy().copy() for k, v in b.views.items()},
           means.numpy().copy(), cam.numpy().copy(), degree))
    dist.destroy_process_group()


def test_factored_sh_bucket_world2():
    """GradBucket(factor_sh=True) over gloo: the reduced fields are the sums / maxima, and views['dsh'] is the sum over the
    ranks of the outer products w(dir(mean, camera_r)) (x) rgb_r -- checked against an independent evaluation of the reference's
    SH backward (computeColorFromSH's gradient through autograd of the forward polynomial)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_factored, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t = lambda a: torch.from_numpy(a)   # noqa: E731
    for name in ("dmeans3D", "dopacity", "dscales", "drot", "dens_sum", "dens_max"):
        a, c = t(res[0][1][name]), t(res[1][1][name])
        total = torch.maximum(a, c) if name == "dens_max" else a + c
        for r in range(world):
            assert torch.equal(t(res[r][2][name]), total), name
    # independent dsh: autograd of colour = sum_k basis_k(dir) * sh_k (the forward polynomial of forward.cu:20-72) w.r.t. sh
    import gof_dp
    means = t(res[0][3])
    want = torch.zeros(means.shape[0], 16, 3, dtype=torch.float64)
    for r in range(world):
        cam, degree, rgb = t(res[r][4]).double(), res[r][5], t(res[r][1]["dsh_rgb"]).double()
        sh = torch.zeros(means.shape[0], 16, 3, dtype=torch.float64, requires_grad=True)
        d = means.double() - cam
        d = d / d.norm(dim=1, keepdim=True)
        basis = gof_dp.sh_grad_weights_torch(d, degree)                      # the basis IS d colour / d sh
        colour = (basis[:, :, None] * sh[:, :basis.shape[1], :]).sum(1)
        (g,) = torch.autograd.grad(colour, sh, grad_outputs=rgb)
        want += g
    for r in range(world):
        got = t(res[r][2]["dsh"]).double()
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    assert torch.equal(t(res[0][2]["dsh"]), t(res[1][2]["dsh"]))              # same bits on both ranks
    # degree 2 on rank 1: its coefficients 9..15 carry rank 0's contribution only
    assert float(t(res[0][2]["dsh"])[:, 9:, :].abs().sum()) > 0


def test_sh_basis_matches_reference_forward_polynomial():
    """gof_dp.sh_grad_weights_torch is the SH basis of the reference's forward (forward.cu:20-72): evaluated against the
    closed-form real spherical harmonics constants at a few directions."""
    import gof_dp
    d = torch.tensor([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [0.6, 0.0, 0.8]], dtype=torch.float64)
    w = gof_dp.sh_grad_weights_torch(d, 3)
    assert w.shape == (3, 16)
    assert torch.allclose(w[:, 0], torch.full((3,), 0.28209479177387814, dtype=torch.float64))
    assert torch.allclose(w[0, 1:4], torch.tensor([0.0, 0.4886025119029199, 0.0], dtype=torch.float64))       # -C1 y, C1 z, -C1 x at +z
    assert torch.allclose(w[1, 1:4], torch.tensor([0.0, 0.0, -0.4886025119029199], dtype=torch.float64))
    assert abs(float(w[0, 6]) - 0.31539156525252005 * 2.0) < 1e-12                                                # C2_2 (2zz - xx - yy) at +z
    assert abs(float(w[2, 7]) - (-1.0925484305920792 * 0.48)) < 1e-12                                             # C2_3 xz
