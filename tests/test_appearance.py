"""CPU: gof_appearance (AppearanceNetwork, l1_loss_appearance) against the reference's OWN code: its AppearanceNetwork class
(scene/appearance_network.py, staged unmodified by baseline/stage_ref.sh) with the same weights, and its L1_loss_appearance
function compiled from the staged train.py text -- values and gradients w.r.t. the image, the embedding and every network
parameter.  The state_dict of one loads into the other.  Skipped when the staged files are absent."""
import importlib.util
import sys
import types

import pytest
import torch

import _refpy
import gof_appearance


@pytest.fixture(scope="module")
def refs():
    path = _refpy.staged("scene", "appearance_network.py")
    if path is None or _refpy.staged("text", "train.py") is None or _refpy.ref_utils("loss_utils") is None:
        pytest.skip("staged reference Python absent (needs /root/reference at build time)")
    spec = importlib.util.spec_from_file_location("gof_ref_appearance_network", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    glb = {"torch": torch, "l1_loss": _refpy.ref_utils("loss_utils").l1_loss}
    return mod.AppearanceNetwork, _refpy.ref_function("train.py", "L1_loss_appearance", glb)


@pytest.mark.parametrize("H,W", [(96, 160), (70, 101)])
def test_network_and_loss_equal_reference(refs, H, W):
    RefNet, ref_loss = refs
    torch.manual_seed(0)
    ref_net = RefNet(67, 3)
    net = gof_appearance.AppearanceNetwork(67, 3)
    net.load_state_dict(ref_net.state_dict())                     # same parameter names / shapes
    assert [k for k, _ in net.named_parameters()] == [k for k, _ in ref_net.named_parameters()]
    emb_table = torch.randn(8, 64) * 1e-2
    image = torch.rand(3, H, W)
    gt = torch.rand(3, H, W)

    def run(loss_fn, network, via_stub):
        img = image.clone().requires_grad_(True)
        table = emb_table.clone().requires_grad_(True)
        for p in network.parameters():
            p.grad = None
        if via_stub:
            g = types.SimpleNamespace(get_apperance_embedding=lambda idx: table[idx], appearance_network=network)
            loss = loss_fn(img, gt, g, 5)
        else:
            loss = loss_fn(img, gt, network, table[5])
        loss.backward()
        return loss.detach(), img.grad, table.grad, [p.grad.clone() for p in network.parameters()]

    lr, gir, gtr, gpr = run(ref_loss, ref_net, True)
    lo, gio, gto, gpo = run(gof_appearance.l1_loss_appearance, net, False)
    assert torch.allclose(lo, lr, rtol=1e-6, atol=1e-7)
    assert torch.allclose(gio, gir, rtol=1e-5, atol=1e-8) and torch.allclose(gto, gtr, rtol=1e-5, atol=1e-8)
    for a, b in zip(gpo, gpr):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-7)
    # the transformed image (evaluation path, train.py:200)
    g = types.SimpleNamespace(get_apperance_embedding=lambda idx: emb_table[idx], appearance_network=ref_net)
    with torch.no_grad():
        tr = ref_loss(image, gt, g, 2, return_transformed_image=True)
        to = gof_appearance.l1_loss_appearance(image, gt, net, emb_table[2], return_transformed_image=True)
    assert torch.allclose(to, tr, rtol=1e-6, atol=1e-7)


def test_flat_gradient_packing_round_trip():
    net = gof_appearance.AppearanceNetwork(67, 3)
    n = gof_appearance.appearance_numel(net)
    assert n == sum(p.numel() for p in net.parameters()) + 64
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(i + 1))
    flat = torch.zeros(n)
    gof_appearance.appearance_grads_flat(net, torch.arange(64.0), flat)
    want = [p.grad.clone() for p in net.parameters()]
    emb = gof_appearance.load_flat_grads_(net, flat * 2)
    for p, w in zip(net.parameters(), want):
        assert torch.equal(p.grad, 2 * w)
    assert torch.equal(emb, 2 * torch.arange(64.0))
