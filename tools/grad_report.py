import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "oracle", "gaussian-opacity-fields_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import _golden, _util, gof_oracle
from diff_gaussian_rasterization import _C
import test_gpu_golden as tg
dev = torch.device("cuda")
for path in _golden.fixture_paths():
    fx = _golden.load(path); fa = tg._fwd_args(fx, dev)
    R, color, radii, geom, binning, img = _C.rasterize_gaussians(*fa)
    grads = _C.rasterize_gaussians_backward(*_util.bwd_args(fa, radii, geom, R, binning, img, torch.from_numpy(fx["dL_dout"]).to(dev)))
    sc = _golden.oracle_scene(fx); _, _, ost = gof_oracle.forward(sc); od = gof_oracle.backward(sc, ost, fx["dL_dout"])
    omap = dict(dmeans2D="dL_dmean2D", dcolors="dL_dcolors", dopacity="dL_dopacity", dmeans3D="dL_dmean3D", dsh="dL_dsh", dscales="dL_dscale", drot="dL_drot", dv2g="dL_dv2g")
    print(fx["name"])
    for n, g in zip(tg.GRAD_ORDER, grads):
        if n not in omap or (n == "dsh" and fx["colors_precomp"].shape[0] > 0): continue
        o = g.cpu().numpy()
        print("   %-9s ours-ref %.2e  ours-fp64 %.2e  ref-fp64 %.2e  refnoise %.2e" % (n, _golden.relerr(o, fx["grad_"+n])[0], _golden.relerr(o, od[omap[n]])[0], _golden.relerr(fx["grad_"+n], od[omap[n]])[0], float(fx["gradnoise_"+n])))
    # diagnostics: per-column dv2g error and K8 isolation
    dv = grads[8].cpu().numpy(); rv = fx["grad_dv2g"]; ov = od["dL_dv2g"]
    col = lambda a, b: [float("%.1e" % (np.abs(a[:, k].astype(np.float64) - b[:, k]).max() / max(np.abs(b[:, k]).max(), 1e-30))) for k in range(10)]
    print("   dv2g per-column ours-fp64", col(dv, ov)); print("   dv2g per-column ref-fp64 ", col(rv, ov))
    if fx["colors_precomp"].shape[0] == 0:
        k8 = gof_oracle.preprocess_backward(sc, ost["radii"], ost["clamped"], grads[1].cpu().numpy(), dv)
        print("   K8 isolation (fp64 K8 on OUR dv2g vs our K8): dscales %.2e drot %.2e dmeans3D %.2e" % (
            _golden.relerr(grads[6].cpu().numpy(), k8["dL_dscale"])[0], _golden.relerr(grads[7].cpu().numpy(), k8["dL_drot"])[0], _golden.relerr(grads[3].cpu().numpy(), k8["dL_dmean3D"])[0]))
        k8r = gof_oracle.preprocess_backward(sc, ost["radii"], ost["clamped"], fx["grad_dcolors"], rv)
        print("   fp64 K8 on REF dv2g vs fp64 truth: dscales %.2e ; fp64 K8 on OUR dv2g vs truth: dscales %.2e" % (
            _golden.relerr(k8r["dL_dscale"], od["dL_dscale"])[0], _golden.relerr(k8["dL_dscale"], od["dL_dscale"])[0]))
