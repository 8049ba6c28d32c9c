"""GPU: deterministic gradient criterion at BASELINE config 2 size (200 k Gaussians, 800x800).

The reference accumulates 17 float atomics per (pixel, Gaussian) in scheduling order, and its view2gaussian chain rule
multiplies that noise by ~1/scale^2 -- two reference runs differ by percents in dL_dscales / dL_drot / dL_dmeans3D, so
"ours == reference within 1e-4" is not a usable criterion for those.  This test pins them against the fp64
evaluation of the SAME formulas instead (oracle/gof_oracle.c: double accumulation, double chain rule):

    err(ours, fp64)  <=  max(2 * err(reference, fp64), 1e-6)          for every gradient tensor, max-norm and relative L2

with err(reference, fp64) measured in the same run on the live reference extension; without it (oracle/_ref not built) the
absolute bounds of the last column of the committed report apply.  Writes gpurun_out/parity_report.json (copied to profiles/)."""
import json
import os

import numpy as np
import pytest
import torch

import _util
import gof_oracle
import gof_synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["dmeans2D", "dcolors", "dopacity", "dmeans3D", "dcov3D", "dsh", "dscales", "drot", "dv2g"]
OMAP = dict(dmeans2D="dL_dmean2D", dopacity="dL_dopacity", dmeans3D="dL_dmean3D", dsh="dL_dsh", dscales="dL_dscale", drot="dL_drot", dv2g="dL_dv2g")
# absolute fallbacks when the live reference is absent: 3x what round 2 measured for this scene
ABS = dict(dmeans2D=3e-6, dopacity=3e-6, dsh=3e-6, dv2g=3e-6, dmeans3D=3e-2, dscales=3e-2, drot=3e-2)


def _err(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)), float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_gradients_against_fp64_oracle_at_c2():
    from diff_gaussian_rasterization import _C as ours
    ref = _util.load_ref()
    dev = torch.device("cuda")
    cam, gs = gof_synth.make_scene("C2", view=3)
    fa = _util.fwd_args(cam, gs, dev)
    H, W = cam.image_height, cam.image_width
    dL = torch.randn(9, H, W, generator=torch.Generator().manual_seed(77))
    R, color, radii, geom, binning, img = ours.rasterize_gaussians(*fa)
    go = ours.rasterize_gaussians_backward(*_util.bwd_args(fa, radii, geom, R, binning, img, dL.to(dev)))
    gr = None
    if ref is not None:
        Rr, cr, radr, ger, binr, imr = ref.rasterize_gaussians(*fa)
        gr = ref.rasterize_gaussians_backward(*_util.bwd_args(fa, radr, ger, Rr, binr, imr, dL.to(dev)))
        gr2 = ref.rasterize_gaussians_backward(*_util.bwd_args(fa, radr, ger, Rr, binr, imr, dL.to(dev)))
    torch.cuda.synchronize()
    sc = gof_oracle.scene_from_synth(cam, gs)
    _out, oradii, st = gof_oracle.forward(sc)
    assert np.array_equal(radii.cpu().numpy(), oradii)
    od = gof_oracle.backward(sc, st, dL.numpy())
    report, failures = {"config": "C2 view 3: 200000 Gaussians, 800x800, sh_degree 3", "criterion": "err(ours,fp64) <= max(2*err(ref,fp64), 1e-6); (max-norm, relative L2)"}, []
    for i, n in enumerate(NAMES):
        if n not in OMAP:
            continue
        eo = _err(go[i].cpu().numpy(), od[OMAP[n]])
        row = {"ours_vs_fp64": eo}
        if gr is not None:
            er = _err(gr[i].cpu().numpy(), od[OMAP[n]])
            row.update(ref_vs_fp64=er, ref_vs_ref=_err(gr2[i].cpu().numpy(), gr[i].cpu().numpy()), ours_vs_ref=_err(go[i].cpu().numpy(), gr[i].cpu().numpy()))
            ok = eo[0] <= max(2 * er[0], 1e-6) and eo[1] <= max(2 * er[1], 1e-6)
        else:
            ok = eo[0] <= ABS[n]
        row["ok"] = bool(ok)
        report[n] = row
        if not ok:
            failures.append((n, row))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as f:
        json.dump(report, f, indent=1)
    assert not failures, failures
