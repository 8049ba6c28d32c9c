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
