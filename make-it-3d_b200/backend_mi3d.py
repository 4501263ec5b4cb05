"""`_backend` object for the REFERENCE's own `raymarching/raymarching.py`, served by libmi3d.so.

The reference resolves its native module in `get_backend()` (raymarching/raymarching.py:14-25: `import _raymarching as _backend`,
else the JIT build of raymarching/src/*.cu).  A maintainer drops this file next to it and writes

    from mi3d_b200.backend_mi3d import _backend          # instead of:  from .backend import _backend

Every function below has the name, positional argument order and in-place output convention of the pybind11 bindings it replaces
(raymarching/src/bindings.cpp:7-22, prototypes raymarching/src/raymarching.h:7-22): the caller allocates every output tensor, the
call returns None, errors raise.  Only the C ABI of include/mi3d.h is used underneath (ctypes, raw device pointers); the work is
enqueued on torch's CURRENT stream (the reference's kernels use the legacy default stream).

Not provided (unreachable in the reference, main.py:54,105-106): sph_from_ray, composite_sdf_rays_train_forward/backward,
composite_sdf_rays -- calling them raises, loudly.
"""
import ctypes as C

import torch

from . import _lib as L


def _p(t):
    return L.ptr(t)


def _f32(t):
    """the reference wrappers hand fp32 contiguous CUDA tensors to the backend (custom_fwd(cast_inputs=float32) + .contiguous())"""
    if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
        raise L.Mi3dError(f"backend_mi3d expects fp32 contiguous CUDA tensors, got {t.dtype} contiguous={t.is_contiguous()} cuda={t.is_cuda}")
    return t


class _backend:
    # ---- raymarching.h:7 ------------------------------------------------------------------------------------------------------
    @staticmethod
    def near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars):
        L.check(L.lib().mi3d_near_far_from_aabb(_p(_f32(rays_o)), _p(_f32(rays_d)), _p(_f32(aabb)), C.c_uint32(N), C.c_float(min_near),
                                                _p(nears), _p(fars), L.stream()), "near_far_from_aabb")

    # ---- raymarching.h:9-11 ---------------------------------------------------------------------------------------------------
    @staticmethod
    def morton3D(coords, N, indices):
        L.check(L.lib().mi3d_morton3D(_p(coords), C.c_uint32(N), _p(indices), L.stream()), "morton3D")

    @staticmethod
    def morton3D_invert(indices, N, coords):
        L.check(L.lib().mi3d_morton3D_invert(_p(indices), C.c_uint32(N), _p(coords), L.stream()), "morton3D_invert")

    @staticmethod
    def packbits(grid, N, density_thresh, bitfield):
        """N = number of bitfield BYTES (raymarching.py:157 passes C * H^3 // 8)."""
        L.check(L.lib().mi3d_packbits(_p(_f32(grid)), C.c_uint32(N), C.c_float(density_thresh), C.c_void_p(0), _p(bitfield), L.stream()),
                "packbits")

    # ---- raymarching.h:13 -----------------------------------------------------------------------------------------------------
    @staticmethod
    def march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C_, H, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises):
        """Two behavioural differences from raymarching.cu:312-480, both documented in include/mi3d.h: samples are compacted in ray-id
        order (deterministic) instead of atomic-arrival order, and counter[0] counts EMITTED samples."""
        lib = L.lib()
        ws = torch.empty(lib.mi3d_march_rays_train_workspace_bytes(C.c_uint32(N)), dtype=torch.uint8, device=rays_o.device)
        L.check(lib.mi3d_march_rays_train(
            _p(_f32(rays_o)), _p(_f32(rays_d)), _p(grid), C.c_float(bound), C.c_float(dt_gamma), C.c_uint32(max_steps), C.c_uint32(N),
            C.c_uint32(C_), C.c_uint32(H), C.c_uint32(M), _p(_f32(nears)), _p(_f32(fars)), C.c_void_p(0), C.c_float(0.0), C.c_void_p(0),
            C.c_void_p(0), _p(_f32(noises)), C.c_uint64(0), _p(xyzs), _p(dirs), _p(deltas), _p(rays), _p(counter), _p(ws), L.stream()),
            "march_rays_train")

    # ---- raymarching.h:14-15 --------------------------------------------------------------------------------------------------
    @staticmethod
    def composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image):
        L.check(L.lib().mi3d_composite_rays_train_forward(
            _p(_f32(sigmas)), _p(_f32(rgbs)), _p(_f32(deltas)), _p(rays), C.c_uint32(M), C.c_uint32(N), C.c_float(T_thresh), _p(weights_sum),
            _p(depth), _p(image), C.c_void_p(0), C.c_void_p(0), C.c_void_p(0), L.stream()), "composite_rays_train_forward")

    @staticmethod
    def composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N, T_thresh,
                                      grad_sigmas, grad_rgbs):
        """the reference pre-zeroes grad_sigmas / grad_rgbs (raymarching.py:295-296); zero_tail=0 keeps exactly that contract"""
        L.check(L.lib().mi3d_composite_rays_train_backward(
            _p(_f32(grad_weights_sum)), _p(_f32(grad_image)), C.c_void_p(0), _p(_f32(sigmas)), _p(_f32(rgbs)), _p(_f32(deltas)), _p(rays),
            _p(_f32(weights_sum)), _p(_f32(image)), C.c_uint32(M), C.c_uint32(N), C.c_float(T_thresh), C.c_void_p(0), _p(grad_sigmas),
            _p(grad_rgbs), C.c_int(0), L.stream()), "composite_rays_train_backward")

    # ---- raymarching.h:20-21 --------------------------------------------------------------------------------------------------
    @staticmethod
    def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C_, H, grid, nears, fars, xyzs, dirs,
                   deltas, noises):
        L.check(L.lib().mi3d_march_rays(
            C.c_uint32(n_alive), C.c_uint32(n_step), _p(rays_alive), _p(_f32(rays_t)), _p(_f32(rays_o)), _p(_f32(rays_d)), C.c_float(bound),
            C.c_float(dt_gamma), C.c_uint32(max_steps), C.c_uint32(C_), C.c_uint32(H), _p(grid), _p(_f32(nears)), _p(_f32(fars)), _p(xyzs),
            _p(dirs), _p(deltas), _p(_f32(noises)), L.stream()), "march_rays")

    @staticmethod
    def composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, normals, deltas, weights_sum, depth, image, normal):
        L.check(L.lib().mi3d_composite_rays(
            C.c_uint32(n_alive), C.c_uint32(n_step), C.c_float(T_thresh), _p(rays_alive), _p(rays_t), _p(_f32(sigmas)), _p(_f32(rgbs)),
            _p(_f32(normals)), _p(_f32(deltas)), _p(weights_sum), _p(depth), _p(image), _p(normal), L.stream()), "composite_rays")

    # ---- unreachable in the reference (bg_radius = -1, --backbone sdf raises) ---------------------------------------------------
    @staticmethod
    def _unbuilt(name):
        def f(*a, **k):
            raise L.Mi3dError(f"{name} is not built: unreachable in the reference (main.py:54,105-106), see INTEGRATION.md")
        return staticmethod(f)


for _n in ("sph_from_ray", "composite_sdf_rays_train_forward", "composite_sdf_rays_train_backward", "composite_sdf_rays"):
    setattr(_backend, _n, _backend._unbuilt(_n))
