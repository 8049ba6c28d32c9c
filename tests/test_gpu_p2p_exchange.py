"""GPU, needs >= 2 devices (skipped on a one-GPU box; run with `gpurun --gpus 2`): the library's own exchange kernels
(csrc/exchange.cu) through gof_dp.GradBucket -- over NVLink peer memory (enable_peer_exchange) and through the NVSwitch
(enable_nvls_exchange, multimem) -- against NCCL's all-reduce on the same buckets: SUM over the gradient part (two addends:
a+b is commutative, so bit-identical), MAX over the statistics tail; every rank must hold the same bits."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q, mode):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import gof_dp
    P = 100_003                                  # not a multiple of 4: the fields are padded to 256 bytes
    bucket = gof_dp.GradBucket(P, 16, dev)
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    mine = torch.randn(bucket.flat.numel(), generator=g)
    mine[bucket.n_sum:] = mine[bucket.n_sum:].abs()          # the MAX tail holds non-negative statistics
    bucket.flat.copy_(mine)
    ref = bucket.flat.clone()
    dist.all_reduce(ref[:bucket.n_sum])          # NCCL result
    dist.all_reduce(ref[bucket.n_sum:], op=dist.ReduceOp.MAX)
    try:
        bucket.enable_nvls_exchange() if mode == "nvls" else bucket.enable_peer_exchange()
    except Exception as e:   # noqa: BLE001 -- symmetric on all ranks
        q.put((rank, "unavailable: " + str(e)[:200], None))
        dist.barrier(); dist.destroy_process_group()
        return
    assert bucket.exchange == mode
    for it in range(3):                          # repeated use: barriers must separate the rounds
        bucket.flat.copy_(mine)
        bucket.all_reduce()
    torch.cuda.synchronize()
    q.put((rank, bucket.flat.cpu().numpy().copy(), ref.cpu().numpy().copy()))
    bucket.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["p2p", "nvls"])
def test_exchange_kernels_equal_nccl(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one node")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, mode)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    if isinstance(res[0][1], str):
        pytest.skip(f"{mode} exchange {res[0][1]}")
    (_, a0, r0), (_, a1, r1) = res
    np.testing.assert_array_equal(a0.view(np.int32), a1.view(np.int32))     # same bits on both ranks
    np.testing.assert_array_equal(a0.view(np.int32), r0.view(np.int32))     # and NCCL's bits (two addends / exact max)


def _worker_factored(rank, world, port, q, mode):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import _util
    import gof_dp
    import gof_synth
    from diff_gaussian_rasterization import _C
    P, H, W = 50_003, 208, 320
    cam, gs = gof_synth.make_scene(dict(P=P, width=W, height=H, seed=17), view=3 + 7 * rank)     # a different view per rank
    fa = _util.fwd_args(cam, gs, dev)
    R, color, radii, geom, binning, img = _C.rasterize_gaussians(*fa)
    grad = torch.randn(9, H, W, generator=torch.Generator().manual_seed(2)).to(dev)
    bucket = gof_dp.GradBucket(P, 16, dev, factor_sh=True)
    try:
        if mode == "nvls":
            bucket.enable_nvls_exchange()
        elif mode == "p2p":
            bucket.enable_peer_exchange()
    except Exception as e:   # noqa: BLE001 -- symmetric on all ranks
        q.put((rank, "unavailable: " + str(e)[:200], None))
        dist.barrier(); dist.destroy_process_group()
        return
    assert bucket.exchange == mode and bucket.factored and bucket.flat.numel() < 0.45 * gof_dp.GradBucket(P, 16, dev).flat.numel()
    names = ("dmeans3D", "dsh", "dopacity", "dscales", "drot", "dens_sum", "dens_max")
    for it in range(2):                                          # repeated use: the records are rewritten every step
        # one backward leaves the record AND (checks only) this view's own full dL_dsh: the blend kernel's float atomics make
        # two backward runs differ in the last bits, so the reference must come from the same run
        full = torch.empty(P, 16, 3, device=dev)
        views = dict(bucket.views)
        views["_dsh_full"] = full
        _C.rasterize_gaussians_backward(*_util.bwd_args(fa, radii, geom, R, binning, img, grad), _out=views)
        want = {n: (full if n == "dsh" else bucket.views[n]).clone() for n in names}
        for n in names:                                          # the plain exchange: NCCL over the unfactored tensors
            dist.all_reduce(want[n], op=dist.ReduceOp.MAX if n == "dens_max" else dist.ReduceOp.SUM)
        bucket.all_reduce(means3D=fa[1])
    torch.cuda.synchronize()
    out = {n: (bucket.views[n].cpu().numpy().copy(), want[n].cpu().numpy().copy()) for n in names}
    q.put((rank, out, None))
    bucket.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["nccl", "p2p", "nvls"])
def test_factored_exchange_equals_plain_allreduce(mode):
    """GradBucket(factor_sh=True) on two GPUs rendering different views: after all_reduce() every field -- dL_dsh expanded from
    the two views' dL_dRGB records included -- carries the same bits as NCCL's all-reduce of the unfactored 64-float bucket
    (two addends: the sum does not depend on the order), on both ranks, for every exchange mode."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one node")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_factored, args=(r, 2, port, q, mode)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    if isinstance(res[0][1], str):
        pytest.skip(f"{mode} exchange {res[0][1]}")
    for name in res[0][1]:
        a0, r0 = res[0][1][name]
        a1, _ = res[1][1][name]
        assert np.abs(r0).max() > 0, name
        np.testing.assert_array_equal(a0.view(np.int32), a1.view(np.int32), err_msg=name)     # same bits on both ranks
        np.testing.assert_array_equal(a0, r0, err_msg=name)                                   # and the plain all-reduce's values
