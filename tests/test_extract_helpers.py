"""CPU: extraction-side host logic (gof_extract): view-sharded evaluate_alpha over gloo (world 2) equals the serial
loop of extract_mesh.py:17-34 including the colour arg-min rule; the bisection converges on an analytic level set."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gof_extract


def _fake_integrate(points, view):
    g = torch.Generator().manual_seed(1000 + int(view))
    a = torch.rand(points.shape[0], generator=g)
    a = torch.where(a > 0.7, torch.ones_like(a), a)          # many points unseen by a view keep alpha 1
    a[::7] = 0.25                                            # exact ties between views
    c = torch.rand(points.shape[0], 3, generator=g)
    return a, c


def _serial(points, views, return_color):
    final_alpha = torch.ones(points.shape[0])
    final_color = torch.ones(points.shape[0], 3)
    for v in views:
        a, c = _fake_integrate(points, v)
        if return_color:
            final_color = torch.where((a < final_alpha).reshape(-1, 1), c, final_color)
        final_alpha = torch.min(final_alpha, a)
    return (1 - final_alpha, final_color) if return_color else 1 - final_alpha


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pts = torch.zeros(500, 3)
    alpha, color = gof_extract.evaluate_alpha(pts, range(9), _fake_integrate, return_color=True)
    alpha2 = gof_extract.evaluate_alpha(pts, range(9), _fake_integrate)
    q.put((rank, alpha.numpy().copy(), color.numpy().copy(), alpha2.numpy().copy()))
    dist.destroy_process_group()


def test_view_sharded_evaluate_alpha_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_a, want_c = _serial(torch.zeros(500, 3), range(9), True)
    for _, a, c, a2 in res:
        assert torch.equal(torch.from_numpy(a), want_a)
        assert torch.equal(torch.from_numpy(a2), want_a)
        assert torch.equal(torch.from_numpy(c), want_c)


def test_single_process_matches_serial():
    pts = torch.zeros(300, 3)
    a, c = gof_extract.evaluate_alpha(pts, range(5), _fake_integrate, return_color=True)
    wa, wc = _serial(pts, range(5), True)
    assert torch.equal(a, wa) and torch.equal(c, wc)


def test_binary_search_converges_on_sphere():
    g = torch.Generator().manual_seed(0)
    inner = torch.nn.functional.normalize(torch.randn(200, 3, generator=g), dim=1) * 0.3
    outer = torch.nn.functional.normalize(torch.randn(200, 3, generator=g), dim=1) * 1.2
    alpha = lambda p: (p.norm(dim=1) < 0.7).float()          # occupancy: 1 inside the sphere of radius 0.7
    end_points = torch.stack([inner, outer], dim=1)
    end_sdf = torch.stack([alpha(inner) - 0.5, alpha(outer) - 0.5], dim=1).reshape(-1, 2, 1)
    pts = gof_extract.binary_search(end_points, end_sdf, alpha, n_steps=8)
    # each step halves the bracket: |r - 0.7| <= |outer - inner| / 2^9
    assert float((pts.norm(dim=1) - 0.7).abs().max()) < 1.5 / 2 ** 8
