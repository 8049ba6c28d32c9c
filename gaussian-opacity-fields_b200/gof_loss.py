"""Per-view training loss of the reference's train.py:151-188 as one fused CUDA pass over the rasterizer's 9-channel
output (csrc/view_loss.cu through the C ABI `gof_view_loss`) -- SURVEY.md section 8(f) rank 1, a CALLER of the rasterizer:

    loss = (1 - lambda_dssim) * L1(rgb, gt) + lambda_dssim * (1 - SSIM(rgb, gt))
           + lambda_depth_normal * mean(1 - n_world . depth_to_normal(depth)) + lambda_distortion * mean(distortion)

    loss, terms = view_loss(rendering, gt_image, viewpoint_cam.world_view_transform, tanfovx, tanfovy,
                            lambda_dssim=0.2, lambda_depth_normal=0.05, lambda_distortion=100.0)
    loss.backward()            # d loss / d rendering comes from the same pass

`terms` = tensor (L1, SSIM, normal-consistency loss, distortion loss, total) for logging.  CUDA tensors only.
"""
import ctypes

import torch

from diff_gaussian_rasterization import _C

_lib = _C._lib
_lib.gof_view_loss_scratch_bytes.restype = ctypes.c_size_t
_lib.gof_view_loss_scratch_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
_lib.gof_view_loss.restype = ctypes.c_int
_lib.gof_view_loss.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float,
                               ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p,
                               ctypes.c_void_p, ctypes.c_void_p]


def _run(rendering, gt, R9, fx, fy, lam, lam_dn, lam_dist, need_grad):
    if not (rendering.is_cuda and gt.is_cuda):
        raise RuntimeError("gof_b200 view_loss: CUDA tensors required (no CPU path)")
    if rendering.dim() != 3 or rendering.shape[0] != 9 or gt.shape != (3,) + tuple(rendering.shape[1:]):
        raise RuntimeError("view_loss: rendering must be (9,H,W) and gt (3,H,W)")
    r, g = rendering.detach().contiguous().float(), gt.detach().contiguous().float()
    H, W = int(r.shape[1]), int(r.shape[2])
    dev = r.device
    terms = torch.empty(5, dtype=torch.float32, device=dev)
    grad = torch.empty_like(r) if need_grad else None
    scratch = torch.empty(int(_lib.gof_view_loss_scratch_bytes(W, H)), dtype=torch.uint8, device=dev)
    Rh = (ctypes.c_float * 9)(*[float(x) for x in R9])
    with torch.cuda.device(dev):
        _C._check(_lib.gof_view_loss(W, H, r.data_ptr(), g.data_ptr(), Rh, fx, fy, lam, lam_dn, lam_dist, terms.data_ptr(),
                                     grad.data_ptr() if need_grad else None, scratch.data_ptr(), _C._stream()))
    return terms, grad


class _ViewLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rendering, gt, R9, fx, fy, lam, lam_dn, lam_dist):
        terms, grad = _run(rendering, gt, R9, fx, fy, lam, lam_dn, lam_dist, rendering.requires_grad)
        ctx.save_for_backward(grad) if grad is not None else None
        ctx.has_grad = grad is not None
        ctx.mark_non_differentiable(terms)
        return terms[4].clone(), terms

    @staticmethod
    def backward(ctx, g_loss, _g_terms):
        if not ctx.has_grad:
            return (None,) * 8
        (grad,) = ctx.saved_tensors
        return grad * g_loss, None, None, None, None, None, None, None


def camera_rotation(world_view_transform):
    """Row-major camera-to-world rotation as 9 Python floats: ((world_view_transform^T)^-1)[:3,:3] (train.py:178)."""
    c2w = torch.linalg.inv(world_view_transform.detach().double().cpu().t())
    return [float(x) for x in c2w[:3, :3].reshape(-1)]


def view_loss(rendering, gt_image, world_view_transform, tanfovx, tanfovy, lambda_dssim=0.2, lambda_depth_normal=0.05,
              lambda_distortion=100.0, rotation=None):
    """Returns (loss, terms).  `rotation` = camera_rotation(world_view_transform) may be passed to avoid the small
    device->host copy per call (cameras are static during training)."""
    H, W = int(rendering.shape[1]), int(rendering.shape[2])
    R9 = rotation if rotation is not None else camera_rotation(world_view_transform)
    return _ViewLoss.apply(rendering, gt_image, R9, W / (2.0 * float(tanfovx)), H / (2.0 * float(tanfovy)), float(lambda_dssim),
                           float(lambda_depth_normal), float(lambda_distortion))
