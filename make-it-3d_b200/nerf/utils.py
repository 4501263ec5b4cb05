"""Row R0 of the hot path: ray generation with the reference's call surface (`get_rays`, nerf/utils.py:51-116), computed by
the `k_get_rays` kernel of libmi3d.so (csrc/raymarch.cu) instead of ~10 small torch ops.  The training path does not even
materialise the rays: `NeRFRenderer.render(..., poses=, intrinsics=, H=, W=)` generates them inside the march kernel
(`mi3d_march_rays_train_cam`).  `safe_normalize` keeps the reference's name for the one host-side use (the light direction)."""
import ctypes as C

import torch

from .. import _lib as L


def safe_normalize(x, eps=1e-20):
    """x / sqrt(clamp(|x|^2, eps, 1e32)) -- the reference's guard against zero-length vectors (nerf/utils.py:47-48)"""
    sq = (x * x).sum(-1, keepdim=True)
    return x / sq.clamp(eps, 1e32).sqrt()


def camera_table(poses, intrinsics, device=None):
    """[B,4,4] cam2world + (fx, fy, cx, cy) -> the [B,16] fp32 device table mi3d_raygen reads (pose rows 0..2, then intrinsics)."""
    poses = torch.as_tensor(poses, dtype=torch.float32)
    if poses.dim() == 2:
        poses = poses[None]
    device = device or poses.device
    B = poses.shape[0]
    intr = torch.as_tensor(intrinsics, dtype=torch.float32).reshape(-1, 4)
    if intr.shape[0] == 1 and B > 1:
        intr = intr.expand(B, 4)
    return torch.cat([poses[:, :3, :].reshape(B, 12).to(device), intr.to(device)], dim=1).contiguous()


def raygen_struct(cams, H, W, rays_per_view=None, stride=1, phase=0):
    rg = L.RayGen()
    rg.cams = cams.data_ptr(); rg.n_views = cams.shape[0]; rg.H = H; rg.W = W
    rg.rays_per_view = H * W if rays_per_view is None else rays_per_view
    rg.pixel_stride = stride; rg.pixel_phase = phase
    rg._keep = cams                 # the struct only holds the raw pointer: keep the table alive as long as the struct
    return rg


@torch.no_grad()
def get_rays(poses, intrinsics, H, W, N=-1, error_map=None):
    """poses [B,4,4] cam2world (CUDA), intrinsics (fx, fy, cx, cy) -> {'rays_o', 'rays_d' [B,HW,3], 'depth_scale' [B,HW], 'inds'}.
    Only N = -1 (every pixel) exists: it is the one mode the training loader uses (nerf/provider.py:297)."""
    if N > 0 or error_map is not None:
        raise NotImplementedError("ray sub-sampling is not on the hot path (SURVEY.md 8a-R0)")
    L.require_cuda(poses)
    B = poses.shape[0]
    cams = camera_table(poses, intrinsics, poses.device)
    rg = raygen_struct(cams, H, W)
    n = B * H * W
    rays_o = torch.empty(B, H * W, 3, dtype=torch.float32, device=poses.device)
    rays_d = torch.empty_like(rays_o)
    scale = torch.empty(B, H * W, dtype=torch.float32, device=poses.device)
    L.check(L.lib().mi3d_get_rays(C.byref(rg), C.c_uint32(n), L.ptr(rays_o), L.ptr(rays_d), L.ptr(scale), L.stream()), "get_rays")
    return {'rays_o': rays_o, 'rays_d': rays_d, 'depth_scale': scale,
            'inds': torch.arange(H * W, device=poses.device).expand([B, H * W])}
