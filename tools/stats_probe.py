import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _util, gof_synth
from diff_gaussian_rasterization import _C as ours
dev = torch.device("cuda")
for name in sys.argv[1:] or ["C3"]:
    cam, gs = gof_synth.make_scene(name, view=1)
    fa = _util.fwd_args(cam, gs, dev)
    grad = torch.randn(9, cam.image_height, cam.image_width, device=dev)
    R, color, radii, geom, binning, img = ours.rasterize_gaussians(*fa)
    ours.rasterize_gaussians_backward(*_util.bwd_args(fa, radii, geom, R, binning, img, grad))
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 8)()
    ours._lib.gof_stats_read(buf)
    v = list(buf)
    print(name, dict(R=R, warp_visits=v[0], lane_evals=v[1], pass_fast_reject=v[2], contributing=v[3], warp_anyhit=v[4], used_sum=v[5], list_sum=v[6]))
    print("  per (tile,G) instance: warp visits %.2f of 8; lanes eval/visit %.1f; pass/eval %.3f; contrib/pass %.3f; anyhit/visit %.3f; contrib/anyhit %.2f; used/list %.3f" % (
        v[0]/max(v[6],1), v[1]/max(v[0],1), v[2]/max(v[1],1), v[3]/max(v[2],1), v[4]/max(v[0],1), v[3]/max(v[4],1), v[5]/max(v[6],1)))
