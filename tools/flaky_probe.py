"""Fresh-process determinism probe: render scene 'a' of tests/_cull_probe.py several times and diff everything."""
import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _util, gof_synth
from diff_gaussian_rasterization import _C
dev = torch.device("cuda")
cfg = dict(P=40000, width=640, height=360, seed=3)
cam = gof_synth.make_camera(cfg["width"], cfg["height"], view=2, radius=4.0)
gs = gof_synth.make_gaussians(cfg["P"], cfg["seed"], cam.focal_x, sigma_px=2.0)
g = torch.Generator().manual_seed(cfg["seed"] + 50)
an = torch.exp(torch.rand(cfg["P"], 3, generator=g) * np.log(1.0 / 0.3) + np.log(0.3))
gs["scales"] = (gs["scales"].max(dim=1, keepdim=True).values * an).contiguous()
fa = _util.fwd_args(cam, gs, dev)
ref = None
for it in range(6):
    R, color, radii, geom, binning, img = _C.rasterize_gaussians(*fa)
    st = _C.export_state(cfg["P"], cfg["width"], cfg["height"], R, geom, binning, img, radii)
    cur = {"color": color.cpu().numpy(), "radii": radii.cpu().numpy()}
    cur.update({k: v.cpu().numpy() for k, v in st.items()})
    if ref is None:
        ref = cur
        import zlib
        print("it0 R", R, "nan", int(np.isnan(cur["color"]).sum()), "crc", {k: zlib.crc32(np.ascontiguousarray(v).tobytes()) & 0xffff for k, v in cur.items()}, flush=True)
    else:
        out = []
        for k in ref:
            a, b = ref[k], cur[k]
            if a.shape != b.shape:
                out.append(f"{k}: shape {a.shape} vs {b.shape}")
            else:
                n = int((a.view(np.uint8) != b.view(np.uint8)).sum()) if a.dtype != np.uint8 else int((a != b).sum())
                if n: out.append(f"{k}: {n} bytes differ")
        print(f"it{it} R {R}:", "; ".join(out) if out else "identical to it0", flush=True)
    del geom, binning, img, st
