"""Level-set extraction helpers around GaussianRasterizer.integrate -- the callers of the integrate path in the
reference's extract_mesh.py, restated with view sharding over the GPUs of one box.

* evaluate_alpha    == evaluage_alpha (extract_mesh.py:17-34): alpha = 1 - min over views of the integrated opacity,
                       optionally the colour of the arg-min view.  `min` is associative, so with torch.distributed
                       initialised every rank processes views[rank::world] and the partial minima are merged with one
                       all_reduce(MIN) (colour: the lowest view index attaining the minimum wins, like the reference's
                       strict `<` update in view order).
* binary_search     == the 8-step bisection of extract_mesh.py:88-102 on the edge endpoints returned by
                       gof_tetmesh.marching_tetrahedra.
* make_integrate_fn == gaussian_renderer.integrate (gaussian_renderer/__init__.py:118-218) for plain tensors.
"""
import torch
import torch.distributed as dist


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


@torch.no_grad()
def evaluate_alpha(points, views, integrate_fn, return_color=False, group=None):
    """integrate_fn(points, view) -> (alpha_integrated [N], color_integrated [N,3])."""
    rank, world = _world(group)
    n, dev = points.shape[0], points.device
    final_alpha = torch.ones(n, dtype=torch.float32, device=dev)
    final_color = torch.ones(n, 3, dtype=torch.float32, device=dev) if return_color else None
    best_view = torch.full((n,), 2 ** 30, dtype=torch.int32, device=dev) if return_color else None
    views = list(views)
    for vi in range(rank, len(views), world):
        alpha_integrated, color_integrated = integrate_fn(points, views[vi])
        if return_color:
            better = alpha_integrated < final_alpha
            final_color = torch.where(better.reshape(-1, 1), color_integrated, final_color)
            best_view = torch.where(better, torch.full_like(best_view, vi), best_view)
        final_alpha = torch.min(final_alpha, alpha_integrated)
    if world > 1:
        local_alpha = final_alpha.clone()
        dist.all_reduce(final_alpha, op=dist.ReduceOp.MIN, group=group)
        if return_color:
            # the winner is the lowest view index whose alpha equals the global minimum (and is < 1, the initial value)
            cand = torch.where((local_alpha == final_alpha) & (local_alpha < 1.0), best_view, torch.full_like(best_view, 2 ** 30))
            win = cand.clone()
            dist.all_reduce(win, op=dist.ReduceOp.MIN, group=group)
            mine = (cand == win) & (win < 2 ** 30)
            contrib = torch.where(mine.reshape(-1, 1), final_color, torch.zeros_like(final_color))
            dist.all_reduce(contrib, op=dist.ReduceOp.SUM, group=group)
            final_color = torch.where((win < 2 ** 30).reshape(-1, 1), contrib, torch.ones_like(contrib))
    alpha = 1 - final_alpha
    return (alpha, final_color) if return_color else alpha


@torch.no_grad()
def binary_search(end_points, end_sdf, eval_alpha, n_steps=8):
    """extract_mesh.py:73-102.  end_points (E,2,3), end_sdf (E,2,1) from marching_tetrahedra; eval_alpha(points)->alpha.
    Returns the refined vertex positions (E,3)."""
    left_points, right_points = end_points[:, 0, :].clone(), end_points[:, 1, :].clone()
    left_sdf, right_sdf = end_sdf[:, 0, :].clone(), end_sdf[:, 1, :].clone()
    points = (left_points + right_points) / 2.
    for _ in range(n_steps):
        mid_points = (left_points + right_points) / 2
        mid_sdf = (eval_alpha(mid_points) - 0.5).reshape(-1, 1)
        ind_low = ((mid_sdf < 0) & (left_sdf < 0)) | ((mid_sdf > 0) & (left_sdf > 0))
        left_sdf[ind_low] = mid_sdf[ind_low]
        right_sdf[~ind_low] = mid_sdf[~ind_low]
        left_points[ind_low.flatten()] = mid_points[ind_low.flatten()]
        right_points[~ind_low.flatten()] = mid_points[~ind_low.flatten()]
        points = (left_points + right_points) / 2
    return points


def make_integrate_fn(means3D, opacities, scales, rotations, shs, sh_degree, settings_for_view):
    """settings_for_view(view) -> GaussianRasterizationSettings.  Returns integrate_fn for evaluate_alpha."""
    from diff_gaussian_rasterization import GaussianRasterizer

    def fn(points, view):
        rs = settings_for_view(view)
        _, alpha_integrated, color_integrated, _ = GaussianRasterizer(rs).integrate(
            points3D=points, means3D=means3D, means2D=torch.zeros_like(means3D), opacities=opacities, shs=shs, scales=scales,
            rotations=rotations)
        return alpha_integrated, color_integrated
    return fn
