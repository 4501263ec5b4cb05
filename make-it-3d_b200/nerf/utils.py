"""Host-side helpers with the reference's names (nerf/utils.py:47-116): safe_normalize, get_rays (row R0)."""
import torch


def safe_normalize(x, eps=1e-20):
    """nerf/utils.py:47-48"""
    return x / torch.sqrt(torch.clamp(torch.sum(x * x, -1, keepdim=True), min=eps, max=1e32))


@torch.no_grad()
def get_rays(poses, intrinsics, H, W, N=-1, error_map=None):
    """Pixel-centre pinhole rays, nerf/utils.py:51-116 (N=-1: every pixel; the only mode training uses, provider.py:297).
    poses [B,4,4] cam2world, intrinsics (fx, fy, cx, cy) -> {'rays_o','rays_d' [B,HW,3], 'depth_scale' [B,HW]}"""
    if N > 0 or error_map is not None:
        raise NotImplementedError("ray sub-sampling is not on the hot path (SURVEY.md 8a-R0)")
    device = poses.device
    B = poses.shape[0]
    fx, fy, cx, cy = intrinsics
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=device), torch.linspace(0, H - 1, H, device=device), indexing='ij')
    i = i.t().reshape([1, H * W]).expand([B, H * W]) + 0.5
    j = j.t().reshape([1, H * W]).expand([B, H * W]) + 0.5
    zs = torch.ones_like(i)
    xs = (i - cx) / fx * zs
    ys = (j - cy) / fy * zs
    directions = torch.stack((xs, ys, zs), dim=-1)
    scale = 1 / directions.pow(2).sum(-1).pow(0.5)
    directions = safe_normalize(directions)
    rays_d = directions @ poses[:, :3, :3].transpose(-1, -2)
    rays_o = poses[..., :3, 3][..., None, :].expand_as(rays_d)
    return {'rays_o': rays_o, 'rays_d': rays_d, 'depth_scale': scale, 'inds': torch.arange(H * W, device=device).expand([B, H * W])}
