#!/usr/bin/env python
"""Developer diagnostic (1 GPU): the factored SH gradient, piece by piece -- K8's record against its own full dL_dsh, the expansion
kernel against the torch expansion, for one and three views."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "gaussian-opacity-fields_b200"))
import _util  # noqa: E402
import gof_dp  # noqa: E402
import gof_synth  # noqa: E402
from diff_gaussian_rasterization import _C  # noqa: E402

dev = torch.device("cuda")
P, H, W = 30_011, 208, 320
plane = (P + 63) // 64 * 64
slot = gof_dp.SH_SLOT_HEADER + 3 * plane
records = torch.zeros(3 * slot, device=dev)
grad = torch.randn(9, H, W, generator=torch.Generator().manual_seed(2)).to(dev)
fulls, means = [], None
for i, view in enumerate((4, 11, 23)):
    cam, gs = gof_synth.make_scene(dict(P=P, width=W, height=H, seed=17), view=view)
    fa = _util.fwd_args(cam, gs, dev)
    R, color, radii, geom, binning, img = _C.rasterize_gaussians(*fa)
    rec = records[i * slot:(i + 1) * slot]
    full = torch.full((P, 16, 3), float("nan"), device=dev)
    out = {"sh_hdr": rec[:64], "dsh_rgb": rec[64:].view(3, plane), "_dsh_full": full}
    fact = _C.rasterize_gaussians_backward(*_util.bwd_args(fa, radii, geom, R, binning, img, grad), _out=out)
    torch.cuda.synchronize()
    means = fa[1]
    fulls.append(full)
    print(f"view {view}: hdr {rec[:4].tolist()} campos {fa[19].tolist()} visible {int((radii > 0).sum())} "
          f"nan in full {int(torch.isnan(full).sum())} |rgb| {float(out['dsh_rgb'].abs().sum()):.4g} |full| {float(full.abs().sum()):.4g}")
    one = gof_dp.sh_grad_from_views_torch(means, [rec], P, 16)
    print("   torch expansion of this record vs K8's own dL_dsh: rel", _util.rel_err(one, full))
    got = torch.full((P, 16, 3), float("nan"), device=dev)
    ptrs = (ctypes.c_void_p * 1)(rec.data_ptr())
    _C._check(_C._lib.gof_sh_grad_from_views(P, 16, 1, means.data_ptr(), ptrs, got.data_ptr(), _C._stream()))
    torch.cuda.synchronize()
    print("   kernel expansion of this record vs K8's own dL_dsh: rel", _util.rel_err(got, full), "equal", bool(torch.equal(got, full)),
          "mismatching elements", int((got != full).sum()))
    bad = (got != full).nonzero()
    if len(bad):
        g, k, c = [int(t) for t in bad[0]]
        print("   first mismatch", (g, k, c), float(got[g, k, c]), float(full[g, k, c]), "rgb", out["dsh_rgb"][:, g].tolist(), "radius", int(radii[g]))
want = (fulls[0] + fulls[1]) + fulls[2]
got = torch.full((P, 16, 3), float("nan"), device=dev)
ptrs = (ctypes.c_void_p * 3)(*[records.data_ptr() + 4 * i * slot for i in range(3)])
_C._check(_C._lib.gof_sh_grad_from_views(P, 16, 3, means.data_ptr(), ptrs, got.data_ptr(), _C._stream()))
torch.cuda.synchronize()
print("three views: rel", _util.rel_err(got, want), "equal", bool(torch.equal(got, want)), "mismatching", int((got != want).sum()))
