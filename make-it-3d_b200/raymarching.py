"""B1 seam: drop-in for the reference's `raymarching` Python module (raymarching/raymarching.py:31-470).

Same function names, argument order and return values; each is a torch.autograd.Function over the C ABI of
libmi3d.so (include/mi3d.h) instead of the pybind `_raymarching` module.  Differences, all deliberate:
  * kernels run on torch's current stream (the reference launches on the legacy default stream);
  * march_rays_train compacts samples in ray-id order (deterministic) and does not zero-fill 268 MB per step;
    it still performs the D2H read of the sample count, because this B1 signature returns sliced tensors.
    The sync-free path is `nerf.renderer.NeRFRenderer.run_cuda` (fused, device-side counts);
  * no `torch.cuda.empty_cache()` per step.
"""
import ctypes as C

import torch
from torch.autograd import Function

from . import _lib as L


def _as_cuda_f32(t):
    if not t.is_cuda:
        t = t.cuda()
    return L.f32c(t)


class _near_far_from_aabb(Function):
    @staticmethod
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        """raymarching.py:34-59 -> mi3d_near_far_from_aabb"""
        rays_o = _as_cuda_f32(rays_o).view(-1, 3)
        rays_d = _as_cuda_f32(rays_d).view(-1, 3)
        aabb = _as_cuda_f32(aabb)
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=torch.float32, device=rays_o.device)
        fars = torch.empty(N, dtype=torch.float32, device=rays_o.device)
        L.check(L.lib().mi3d_near_far_from_aabb(L.ptr(rays_o), L.ptr(rays_d), L.ptr(aabb), C.c_uint32(N), C.c_float(min_near),
                                                L.ptr(nears), L.ptr(fars), L.stream()), "near_far_from_aabb")
        return nears, fars


near_far_from_aabb = _near_far_from_aabb.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        """raymarching.py:97-114"""
        if not coords.is_cuda:
            coords = coords.cuda()
        coords = coords.int().contiguous()
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        L.check(L.lib().mi3d_morton3D(L.ptr(coords), C.c_uint32(N), L.ptr(indices), L.stream()), "morton3D")
        return indices


morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        """raymarching.py:120-136"""
        if not indices.is_cuda:
            indices = indices.cuda()
        indices = indices.int().contiguous()
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        L.check(L.lib().mi3d_morton3D_invert(L.ptr(indices), C.c_uint32(N), L.ptr(coords), L.stream()), "morton3D_invert")
        return coords


morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    def forward(ctx, grid, thresh, bitfield=None):
        """raymarching.py:144-165; grid [C, H^3] float, bitfield uint8 [C*H^3/8]"""
        grid = _as_cuda_f32(grid)
        C_, H3 = grid.shape[0], grid.shape[1]
        N = C_ * H3 // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        L.check(L.lib().mi3d_packbits(L.ptr(grid), C.c_uint32(N), C.c_float(float(thresh)), C.c_void_p(0), L.ptr(bitfield), L.stream()),
                "packbits")
        return bitfield


packbits = _packbits.apply


class _march_rays_train(Function):
    @staticmethod
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C_, H, nears, fars, step_counter=None, mean_count=-1,
                perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        """raymarching.py:176-245 -> mi3d_march_rays_train (single launch: count, look-back scan, emit)."""
        rays_o = _as_cuda_f32(rays_o).view(-1, 3)
        rays_d = _as_cuda_f32(rays_d).view(-1, 3)
        if not density_bitfield.is_cuda:
            density_bitfield = density_bitfield.cuda()
        density_bitfield = density_bitfield.contiguous()
        nears, fars = _as_cuda_f32(nears), _as_cuda_f32(fars)
        dev = rays_o.device
        N = rays_o.shape[0]
        M = N * max_steps
        if not force_all_rays and mean_count > 0:
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count
        # rows past the emitted samples must read as zeros for callers that look at the aligned tail
        xyzs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        dirs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        deltas = torch.empty(M, 2, dtype=torch.float32, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
        noises = torch.rand(N, dtype=torch.float32, device=dev) if perturb else torch.zeros(N, dtype=torch.float32, device=dev)
        ws = torch.empty(L.lib().mi3d_march_rays_train_workspace_bytes(C.c_uint32(N)), dtype=torch.uint8, device=dev)
        L.check(L.lib().mi3d_march_rays_train(
            L.ptr(rays_o), L.ptr(rays_d), L.ptr(density_bitfield), C.c_float(bound), C.c_float(dt_gamma), C.c_uint32(max_steps),
            C.c_uint32(N), C.c_uint32(C_), C.c_uint32(H), C.c_uint32(M), L.ptr(nears), L.ptr(fars), C.c_void_p(0), C.c_float(0.0),
            C.c_void_p(0), C.c_void_p(0), L.ptr(noises), C.c_uint64(0), L.ptr(xyzs), L.ptr(dirs), L.ptr(deltas), L.ptr(rays),
            L.ptr(step_counter), L.ptr(ws), L.stream()), "march_rays_train")
        if force_all_rays or mean_count <= 0:
            total = int(step_counter[0].item())          # D2H, inherent to this B1 signature (raymarching.py:236)
            m = total
            if align > 0:
                m += align - m % align
            m = min(m, M)
            used = min(total, M)
            xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
            xyzs[used:].zero_(); dirs[used:].zero_(); deltas[used:].zero_()
        else:
            used = torch.clamp(step_counter[0], max=M)
            idx = torch.arange(M, device=dev)
            pad = (idx >= used)
            xyzs[pad] = 0; dirs[pad] = 0; deltas[pad] = 0
        return xyzs, dirs, deltas, rays


march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        """raymarching.py:253-281"""
        sigmas, rgbs = L.f32c(sigmas), L.f32c(rgbs)
        deltas = L.f32c(deltas)
        L.require_cuda(sigmas, rgbs, deltas, rays)
        M, N = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
        weights_sum = torch.empty(N, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        image = torch.empty(N, 3, dtype=torch.float32, device=dev)
        L.check(L.lib().mi3d_composite_rays_train_forward(
            L.ptr(sigmas), L.ptr(rgbs), L.ptr(deltas), L.ptr(rays), C.c_uint32(M), C.c_uint32(N), C.c_float(T_thresh),
            L.ptr(weights_sum), L.ptr(depth), L.ptr(image), C.c_void_p(0), C.c_void_p(0), C.c_void_p(0), L.stream()),
            "composite_rays_train_forward")
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        ctx.dims = [M, N, T_thresh]
        return weights_sum, depth, image

    @staticmethod
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        """raymarching.py:283-300 (grad_depth is not propagated, like the reference)"""
        grad_weights_sum = L.f32c(grad_weights_sum)
        grad_image = L.f32c(grad_image)
        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        L.check(L.lib().mi3d_composite_rays_train_backward(
            L.ptr(grad_weights_sum), L.ptr(grad_image), C.c_void_p(0), L.ptr(sigmas), L.ptr(rgbs), L.ptr(deltas), L.ptr(rays),
            L.ptr(weights_sum), L.ptr(image), C.c_uint32(M), C.c_uint32(N), C.c_float(T_thresh), C.c_void_p(0),
            L.ptr(grad_sigmas), L.ptr(grad_rgbs), C.c_int(0), L.stream()), "composite_rays_train_backward")
        return grad_sigmas, grad_rgbs, None, None, None


composite_rays_train = _composite_rays_train.apply


class _march_rays(Function):
    @staticmethod
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C_, H, near, far, align=-1,
                perturb=False, dt_gamma=0, max_steps=1024):
        """raymarching.py:368-414"""
        rays_o = _as_cuda_f32(rays_o).view(-1, 3)
        rays_d = _as_cuda_f32(rays_d).view(-1, 3)
        dev = rays_o.device
        M = n_alive * n_step
        if align > 0:
            M += align - (M % align)
        xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
        dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
        deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
        noises = torch.rand(n_alive, dtype=torch.float32, device=dev) if perturb else torch.zeros(n_alive, dtype=torch.float32, device=dev)
        L.check(L.lib().mi3d_march_rays(
            C.c_uint32(n_alive), C.c_uint32(n_step), L.ptr(rays_alive), L.ptr(rays_t), L.ptr(rays_o), L.ptr(rays_d), C.c_float(bound),
            C.c_float(dt_gamma), C.c_uint32(max_steps), C.c_uint32(C_), C.c_uint32(H), L.ptr(density_bitfield.contiguous()),
            L.ptr(near), L.ptr(far), L.ptr(xyzs), L.ptr(dirs), L.ptr(deltas), L.ptr(noises), L.stream()), "march_rays")
        return xyzs, dirs, deltas


march_rays = _march_rays.apply


class _composite_rays(Function):
    @staticmethod
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, normals, deltas, weights_sum, depth, image, normal,
                T_thresh=1e-2):
        """raymarching.py:422-447 (in-place on rays_alive, rays_t, weights_sum, depth, image, normal)"""
        sigmas, rgbs, normals = L.f32c(sigmas), L.f32c(rgbs), L.f32c(normals)
        L.check(L.lib().mi3d_composite_rays(
            C.c_uint32(n_alive), C.c_uint32(n_step), C.c_float(T_thresh), L.ptr(rays_alive), L.ptr(rays_t), L.ptr(sigmas),
            L.ptr(rgbs), L.ptr(normals), L.ptr(deltas), L.ptr(weights_sum), L.ptr(depth), L.ptr(image), L.ptr(normal),
            L.stream()), "composite_rays")
        return tuple()


composite_rays = _composite_rays.apply
