"""TEST INFRASTRUCTURE -- CPU restatement (numpy, float64) of the reference's per-view training loss and of its gradient
with respect to the 9-channel render:

    train.py:151-188      loss = (1-l)*L1 + l*(1-SSIM) + l_dn * mean(1 - n_world . n_depth) + l_dist * mean(distortion)
    utils/loss_utils.py:17-63   l1_loss, gaussian window (11, sigma 1.5), _ssim (zero-padded depthwise convolutions)
    utils/depth_utils.py:6-35   depths_to_points, depth_to_normal (central differences, border pixels zero)

Pinned by tests/golden/loss_*.npz, which tests/golden/make_golden_loss.py produced by running the reference's own Python
(values and autograd gradients, float32).  Only tests may import this file."""
import math

import numpy as np

C1, C2 = 0.01 ** 2, 0.03 ** 2


def window_1d(size=11, sigma=1.5):
    g = np.array([math.exp(-(x - size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(size)], np.float64)   # loss_utils.py:23-25
    return g / g.sum()


def blur(img, g):
    """Zero-padded 'same' correlation of every channel with outer(g, g) (F.conv2d, padding = size//2, groups = C)."""
    r = len(g) // 2
    C, H, W = img.shape
    p = np.zeros((C, H + 2 * r, W + 2 * r), np.float64)
    p[:, r:r + H, r:r + W] = img
    tmp = sum(g[k] * p[:, :, k:k + W] for k in range(len(g)))            # along x
    return sum(g[k] * tmp[:, k:k + H, :] for k in range(len(g)))         # along y


def ssim_terms(img1, img2, g):
    mu1, mu2 = blur(img1, g), blur(img2, g)
    e11, e22, e12 = blur(img1 * img1, g), blur(img2 * img2, g), blur(img1 * img2, g)
    s11, s22, s12 = e11 - mu1 * mu1, e22 - mu2 * mu2, e12 - mu1 * mu2
    A1, A2 = 2 * mu1 * mu2 + C1, 2 * s12 + C2
    B1, B2 = mu1 * mu1 + mu2 * mu2 + C1, s11 + s22 + C2
    return mu1, mu2, A1, A2, B1, B2


def camera_terms(world_view_transform, tanfovx, tanfovy, W, H):
    c2w = np.linalg.inv(np.asarray(world_view_transform, np.float64).T)       # train.py:178, depth_utils.py:7
    R, ro = c2w[:3, :3], c2w[:3, 3]
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)                          # depth_utils.py:9-10
    xs, ys = np.arange(W) + 0.5, np.arange(H) + 0.5
    k = np.stack([np.broadcast_to((xs - W / 2.0) / fx, (H, W)), np.broadcast_to(((ys - H / 2.0) / fy)[:, None], (H, W)),
                  np.ones((H, W))], axis=-1)                                   # K^-1 (x, y, 1)
    rays_d = k @ R.T                                                           # depth_utils.py:18
    return R, ro, rays_d


def view_loss(render, gt, world_view_transform, tanfovx, tanfovy, lambdas, need_grad=True):
    """Returns dict(loss, Ll1, ssim, depth_normal_loss, distortion_loss, depth_normal[3,H,W], grad[9,H,W])."""
    render, gt = np.asarray(render, np.float64), np.asarray(gt, np.float64)
    lam, lam_dn, lam_dist = (float(x) for x in lambdas)
    _, H, W = render.shape
    N, N3 = H * W, 3 * H * W
    img = render[:3]
    g = window_1d()
    # ---- L1 + SSIM ----
    Ll1 = np.abs(img - gt).mean()
    mu1, mu2, A1, A2, B1, B2 = ssim_terms(img, gt, g)
    smap = (A1 * A2) / (B1 * B2)
    ssim = smap.mean()
    # ---- distortion ----
    dist = render[8].mean()
    # ---- depth -> normal (depth_utils.py:25-35) ----
    R, ro, rays_d = camera_terms(world_view_transform, tanfovx, tanfovy, W, H)
    P = render[6][..., None] * rays_d + ro
    dn = np.zeros((H, W, 3))
    dxv = P[2:, 1:-1] - P[:-2, 1:-1]
    dyv = P[1:-1, 2:] - P[1:-1, :-2]
    c = np.cross(dxv, dyv)
    cl = np.maximum(np.linalg.norm(c, axis=-1, keepdims=True), 1e-12)        # F.normalize eps
    dn[1:-1, 1:-1] = c / cl
    # ---- rendered normal -> world, consistency (train.py:175-183) ----
    n = np.moveaxis(render[3:6], 0, -1)                                       # (H,W,3)
    nl = np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), 1e-12)
    u = n / nl
    nw = u @ R.T
    err = 1.0 - (nw * dn).sum(-1)
    dnl = err.mean()
    loss = (1.0 - lam) * Ll1 + lam * (1.0 - ssim) + lam_dn * dnl + lam_dist * dist
    out = dict(loss=loss, Ll1=Ll1, ssim=ssim, depth_normal_loss=dnl, distortion_loss=dist, depth_normal=np.moveaxis(dn, -1, 0))
    if not need_grad:
        return out
    grad = np.zeros_like(render)
    # L1
    grad[:3] += (1.0 - lam) * np.sign(img - gt) / N3
    # SSIM: map = A1 A2 / (B1 B2) as a function of (mu1, E11, E12); gradient flows back through three blurs
    dm_dmu1 = (2 * mu2 * A2 - 2 * mu2 * A1) / (B1 * B2) - (A1 * A2) * (2 * mu1 * B2 - 2 * mu1 * B1) / (B1 * B2) ** 2
    dm_de11 = -(A1 * A2) / (B1 * B2 * B2)
    dm_de12 = 2 * A1 / (B1 * B2)
    dS = (blur(dm_dmu1, g) + 2 * img * blur(dm_de11, g) + gt * blur(dm_de12, g)) / N3
    grad[:3] += -lam * dS
    # distortion
    grad[8] += lam_dist / N
    # normal consistency: d/d(render normal)
    g_u = -(dn @ R) / N * lam_dn                                              # dL/du = -R^T dn / N
    small = (np.linalg.norm(n, axis=-1, keepdims=True) <= 1e-12)
    g_n = np.where(small, g_u / 1e-12, (g_u - u * (u * g_u).sum(-1, keepdims=True)) / nl)
    grad[3:6] += np.moveaxis(g_n, -1, 0)
    # normal consistency: d/d(depth) through depth_to_normal (interior pixels only)
    g_dn = -nw[1:-1, 1:-1] / N * lam_dn
    d = dn[1:-1, 1:-1]
    big = (np.linalg.norm(c, axis=-1, keepdims=True) > 1e-12)
    g_c = np.where(big, (g_dn - d * (d * g_dn).sum(-1, keepdims=True)) / cl, g_dn / 1e-12)
    g_a = np.cross(dyv, g_c)                                                  # c = a x b: dc.g = da.(b x g) + db.(g x a)
    g_b = np.cross(g_c, dxv)
    gP = np.zeros_like(P)
    gP[2:, 1:-1] += g_a
    gP[:-2, 1:-1] -= g_a
    gP[1:-1, 2:] += g_b
    gP[1:-1, :-2] -= g_b
    grad[6] += (gP * rays_d).sum(-1)
    out["grad"] = grad
    return out
