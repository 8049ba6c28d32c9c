"""GPU: gof_densify.densify_and_prune against the reference's OWN methods (scene/gaussian_model.py: densify_and_prune,
densify_and_clone, densify_and_split, densification_postfix, cat_tensors_to_optimizer, prune_points, _prune_optimizer --
compiled from the staged, unmodified source text into a stub class around a real torch.optim.Adam).  torch.normal is the only
thing replaced on the reference side: it returns std * the same standard-normal samples our kernel is given, so every surviving
parameter row, both Adam moments and the row order can be compared exactly (positions / scalings to float rounding)."""
import types

import pytest
import torch

import _refpy

pytestmark = pytest.mark.gpu
METHODS = ["densify_and_prune", "densify_and_clone", "densify_and_split", "densification_postfix", "cat_tensors_to_optimizer",
           "prune_points", "_prune_optimizer"]
NAMES = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity", "scaling": "_scaling", "rotation": "_rotation"}


def _reference_model(P, dev, gen):
    gu = _refpy.ref_utils("general_utils")
    if gu is None or _refpy.staged("text", "gaussian_model.py") is None:
        pytest.skip("staged reference Python absent (needs /root/reference at build time)")
    feed = {"queue": []}

    class TorchProxy:
        def __getattr__(self, k):
            return getattr(torch, k)

        @staticmethod
        def normal(mean, std):
            return std * feed["queue"].pop(0)

    glb = {"torch": TorchProxy(), "nn": torch.nn, "build_rotation": gu.build_rotation}
    ns = {}
    for m in METHODS:
        exec(_refpy.ref_method_source("gaussian_model.py", "GaussianModel", m), glb, ns)
    Stub = type("GaussianModelStub", (), dict(ns))
    Stub.get_xyz = property(lambda s: s._xyz)
    Stub.get_scaling = property(lambda s: torch.exp(s._scaling))
    Stub.get_opacity = property(lambda s: torch.sigmoid(s._opacity))
    g = Stub()
    g.scaling_inverse_activation = torch.log
    g.percent_dense = 0.01
    r = lambda *shape: torch.randn(*shape, generator=gen).to(dev)
    g._xyz = torch.nn.Parameter(r(P, 3))
    g._features_dc = torch.nn.Parameter(r(P, 1, 3))
    g._features_rest = torch.nn.Parameter(r(P, 15, 3))
    g._opacity = torch.nn.Parameter(r(P, 1) * 2.0)
    g._scaling = torch.nn.Parameter(torch.log(torch.rand(P, 3, generator=gen) * 0.04 + 1e-3).to(dev))
    g._rotation = torch.nn.Parameter(r(P, 4))
    groups = [{"params": [getattr(g, a)], "lr": 1e-3, "name": n} for n, a in NAMES.items()]
    groups.append({"params": [torch.nn.Parameter(r(4, 4))], "lr": 1e-3, "name": "appearance_network"})      # skipped by the reference's loops
    g.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    for grp in g.optimizer.param_groups:                      # one step so that every group has Adam state
        grp["params"][0].grad = torch.randn(grp["params"][0].shape, generator=gen).to(dev)
    g.optimizer.step()
    g.xyz_gradient_accum = (torch.rand(P, 1, generator=gen) * 6e-4).to(dev)
    g.xyz_gradient_accum_abs = (torch.rand(P, 1, generator=gen) * 9e-4).to(dev)
    g.xyz_gradient_accum_abs_max = torch.zeros(P, 1, device=dev)
    g.denom = torch.randint(0, 3, (P, 1), generator=gen).float().to(dev)              # zeros: the NaN -> 0 path
    g.max_radii2D = (torch.rand(P, generator=gen) * 40).to(dev)
    return g, feed


@pytest.mark.parametrize("P,max_screen", [(20_003, 20), (5_000, None), (257, 20)])
def test_densify_and_prune_equals_reference(P, max_screen):
    import gof_densify
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(P)
    g, feed = _reference_model(P, dev, gen)
    params = {n: getattr(g, a).detach().clone() for n, a in NAMES.items()}
    m = {n: g.optimizer.state[getattr(g, a)]["exp_avg"].clone() for n, a in NAMES.items()}
    v = {n: g.optimizer.state[getattr(g, a)]["exp_avg_sq"].clone() for n, a in NAMES.items()}
    acc, acc_abs, den = g.xyz_gradient_accum.clone(), g.xyz_gradient_accum_abs.clone(), g.denom.clone()
    extent, max_grad, min_op = 1.7, 2e-4, 0.05
    noise = torch.randn(3, P, 3, generator=gen).to(dev)

    ours = gof_densify.densify_and_prune(params, m, v, acc, acc_abs, den, max_grad, min_op, extent, max_screen, noise=noise)

    # the samples the reference's two torch.normal calls will consume: rows of the selected Gaussians, in order
    grads = (acc / den).nan_to_num(nan=0.0, posinf=float("inf")).reshape(-1)
    grads_abs = (acc_abs / den).reshape(-1); grads_abs[grads_abs.isnan()] = 0.0
    grads = (acc / den).reshape(-1); grads[grads.isnan()] = 0.0
    ratio = (grads >= max_grad).float().mean()
    Q = torch.quantile(grads_abs, 1 - ratio)
    sel = (grads >= max_grad) | (grads_abs >= Q)
    smax = torch.exp(params["scaling"]).max(dim=1).values
    clone, split = sel & (smax <= 0.01 * extent), sel & (smax > 0.01 * extent)
    feed["queue"] = [noise[0][clone], torch.cat([noise[1][split], noise[2][split]])]
    g.densify_and_prune(max_grad, min_op, extent, max_screen)
    assert not feed["queue"]

    assert ours.params["xyz"].shape[0] == g._xyz.shape[0] == sum(ours.counts)
    for n, a in NAMES.items():
        ref_p = getattr(g, a).detach()
        st = g.optimizer.state[getattr(g, a)]
        if n in ("xyz", "scaling"):
            assert torch.allclose(ours.params[n], ref_p, rtol=1e-5, atol=1e-6), n
        else:
            assert torch.equal(ours.params[n], ref_p), n
        assert torch.equal(ours.exp_avg[n], st["exp_avg"]), n
        assert torch.equal(ours.exp_avg_sq[n], st["exp_avg_sq"]), n
    assert ours.counts[1] > 0 and ours.counts[2] > 0 and ours.counts[2] == ours.counts[3]
    # the library's own sampler: same decisions, positions within a few sigma of the parents
    own = gof_densify.densify_and_prune(params, m, v, acc, acc_abs, den, max_grad, min_op, extent, max_screen, seed=7)
    assert own.counts == ours.counts and torch.equal(own.src_index, ours.src_index)
    new = own.kind > 0
    parent = params["xyz"][own.src_index.long()]
    dist = (own.params["xyz"] - parent).norm(dim=1)
    assert float(dist[~new].max()) == 0.0
    sig = torch.exp(params["scaling"])[own.src_index.long()].max(dim=1).values
    assert float((dist[new] / sig[new]).max()) < 8.0 and float((dist[new] / sig[new]).mean()) > 0.3
