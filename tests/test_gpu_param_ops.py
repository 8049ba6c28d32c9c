"""GPU: gof_params.activate / adam_step / compute_3d_filter against the
golden vectors generated from the reference's own Python (tests/golden/make_golden_params.py).  The arithmetic itself is
checked on the CPU in test_param_ops_host.py."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FIX = sorted(glob.glob(os.path.join(HERE, "golden", "params_*.npz")))


def _rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[:-4] for p in FIX])
def test_activate_and_adam(path):
    import gof_params
    fx = np.load(path)
    dev = torch.device("cuda")
    t = lambda k: torch.from_numpy(fx[k]).to(dev)
    raw = {k: t("raw" + k).requires_grad_(True) for k in ("_scaling", "_rotation", "_opacity", "_features_dc", "_features_rest")}
    outs = gof_params.activate(raw["_scaling"], raw["_rotation"], raw["_opacity"], t("filter_3D"), raw["_features_dc"], raw["_features_rest"])
    names = ("scales", "rotations", "opacities", "shs")
    for n, o in zip(names, outs):
        assert _rel(o.detach().cpu().numpy(), fx["out_" + n]) < 5e-6, n
    sum((o * t("up_" + n)).sum() for n, o in zip(names, outs)).backward()
    ok = np.linalg.norm(fx["raw_rotation"], axis=1) > 1e-6
    for k in raw:
        g, r = raw[k].grad.cpu().numpy(), fx["grad" + k]
        if k == "_rotation":
            assert _rel(g[ok], r[ok]) < 2e-5 and _rel(g[~ok], r[~ok]) < 1e-5
        else:
            assert _rel(g, r) < 2e-5, k
    p = t("adam_p0").contiguous()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step, g in enumerate(fx["adam_grads"], start=1):
        gof_params.adam_step(p, m, v, torch.from_numpy(g).to(dev).contiguous(), float(fx["adam_lr"]), step)
    assert _rel(m.cpu().numpy(), fx["adam_m"]) < 1e-6 and _rel(v.cpu().numpy(), fx["adam_v"]) < 1e-6
    assert float(np.abs(p.cpu().numpy() - fx["adam_p"]).max()) < 1e-6


def test_compute_3d_filter():
    import gof_params
    fx = np.load(os.path.join(HERE, "golden", "filter3d_a.npz"))
    dev = torch.device("cuda")
    cams = torch.from_numpy(fx["cams"]).to(dev)
    out = gof_params.compute_3d_filter(torch.from_numpy(fx["xyz"]).to(dev), cams, float(fx["cams"][:, 12].max()))
    ref = fx["filter_3D"]
    rel = np.abs(out.cpu().numpy() - ref) / ref
    assert np.quantile(rel, 0.999) < 1e-5 and (rel > 1e-5).sum() <= 3
