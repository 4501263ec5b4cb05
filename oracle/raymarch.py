"""oracle/raymarch.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy-facing ctypes wrapper around oracle/liboracle.so (the C restatement of
raymarching/src/raymarching.cu, see raymarch_oracle.c for per-function file:line citations).
Signatures mirror raymarching/raymarching.py:31-470 minus autograd.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_hashgrid_levels.restype = C.c_uint32
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3), _f32(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().orc_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), C.c_uint32(N), C.c_float(min_near), _p(nears), _p(fars))
    return nears, fars


def morton3D(coords):
    coords = np.ascontiguousarray(coords, dtype=np.int32)
    N = coords.shape[0]
    out = np.empty(N, np.int32)
    lib().orc_morton3D(_p(coords), C.c_uint32(N), _p(out))
    return out


def morton3D_invert(indices):
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    N = indices.shape[0]
    out = np.empty((N, 3), np.int32)
    lib().orc_morton3D_invert(_p(indices), C.c_uint32(N), _p(out))
    return out


def packbits(grid, thresh):
    grid = _f32(grid)
    nbytes = grid.size // 8
    out = np.empty(nbytes, np.uint8)
    lib().orc_packbits(_p(grid), C.c_uint32(nbytes), C.c_float(thresh), _p(out))
    return out


def march_rays_train(rays_o, rays_d, bound, bitfield, Cc, H, nears, fars, noises, dt_gamma=0.0, max_steps=1024,
                     M=None, align=-1):
    """Returns (xyzs[m,3], dirs[m,3], deltas[m,2], rays[N,3], total) with m = total padded like
    raymarching.py:237-241 when align > 0 (note: always adds align - m % align, even when m % align == 0)."""
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    nears, fars, noises = _f32(nears), _f32(fars), _f32(noises)
    N = rays_o.shape[0]
    cap = N * max_steps if M is None else M
    # first a counting-only call would need a second API; the oracle simply allocates worst case lazily
    xyzs = np.zeros((cap, 3), np.float32)
    dirs = np.zeros((cap, 3), np.float32)
    deltas = np.zeros((cap, 2), np.float32)
    rays = np.empty((N, 3), np.int32)
    counter = np.zeros(2, np.int32)
    lib().orc_march_rays_train(_p(rays_o), _p(rays_d), _p(bitfield), C.c_float(bound), C.c_float(dt_gamma),
                               C.c_uint32(max_steps), C.c_uint32(N), C.c_uint32(Cc), C.c_uint32(H), C.c_uint32(cap),
                               _p(nears), _p(fars), _p(xyzs), _p(dirs), _p(deltas), _p(rays), _p(counter), _p(noises))
    m = int(counter[0])
    total = m
    if align > 0:
        m += align - m % align
    m = min(m, cap)
    return xyzs[:m].copy(), dirs[:m].copy(), deltas[:m].copy(), rays, total


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    rays = np.ascontiguousarray(rays, dtype=np.int32)
    M, N = sigmas.shape[0], rays.shape[0]
    ws, depth, image = np.empty(N, np.float32), np.empty(N, np.float32), np.empty((N, 3), np.float32)
    lib().orc_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(deltas), _p(rays), C.c_uint32(M), C.c_uint32(N),
                                           C.c_float(T_thresh), _p(ws), _p(depth), _p(image))
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, T_thresh=1e-4):
    grad_ws, grad_image = _f32(grad_ws), _f32(grad_image)
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    weights_sum, image = _f32(weights_sum), _f32(image)
    rays = np.ascontiguousarray(rays, dtype=np.int32)
    M, N = sigmas.shape[0], rays.shape[0]
    gs, gr = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    lib().orc_composite_rays_train_backward(_p(grad_ws), _p(grad_image), _p(sigmas), _p(rgbs), _p(deltas), _p(rays),
                                            _p(weights_sum), _p(image), C.c_uint32(M), C.c_uint32(N), C.c_float(T_thresh),
                                            _p(gs), _p(gr))
    return gs, gr


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, Cc, H, nears, fars, noises,
               align=-1, dt_gamma=0.0, max_steps=1024):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    rays_alive = np.ascontiguousarray(rays_alive, dtype=np.int32)
    rays_t, nears, fars, noises = _f32(rays_t), _f32(nears), _f32(fars), _f32(noises)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    lib().orc_march_rays(C.c_uint32(n_alive), C.c_uint32(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d),
                         C.c_float(bound), C.c_float(dt_gamma), C.c_uint32(max_steps), C.c_uint32(Cc), C.c_uint32(H),
                         _p(bitfield), _p(nears), _p(fars), _p(xyzs), _p(dirs), _p(deltas), _p(noises))
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, normals, deltas, weights_sum, depth, image, normal,
                   T_thresh=1e-2):
    """In-place on rays_alive, rays_t, weights_sum, depth, image, normal (must be C-contiguous numpy of the right dtype)."""
    sigmas, rgbs, normals, deltas = _f32(sigmas), _f32(rgbs), _f32(normals), _f32(deltas)
    for a, dt in ((rays_alive, np.int32), (rays_t, np.float32), (weights_sum, np.float32), (depth, np.float32),
                  (image, np.float32), (normal, np.float32)):
        assert a.dtype == dt and a.flags.c_contiguous
    lib().orc_composite_rays(C.c_uint32(n_alive), C.c_uint32(n_step), C.c_float(T_thresh), _p(rays_alive), _p(rays_t),
                             _p(sigmas), _p(rgbs), _p(normals), _p(deltas), _p(weights_sum), _p(depth), _p(image), _p(normal))


def hashgrid_levels(L=16, base_res=16, per_level_scale=1.3819128274917603, log2_T=19):
    offsets, sizes, ress = (np.empty(L, np.uint32) for _ in range(3))
    scales = np.empty(L, np.float32)
    total = lib().orc_hashgrid_levels(C.c_uint32(L), C.c_uint32(base_res), C.c_double(per_level_scale), C.c_uint32(log2_T),
                                      _p(offsets), _p(sizes), _p(ress), _p(scales))
    return dict(offsets=offsets, sizes=sizes, ress=ress, scales=scales, total=int(total))


def hashgrid_forward(x, table, levels):
    x, table = _f32(x).reshape(-1, 3), _f32(table)
    E, L = x.shape[0], len(levels["sizes"])
    out = np.empty((E, 2 * L), np.float32)
    lib().orc_hashgrid_forward(_p(x), C.c_uint32(E), _p(table), C.c_uint32(L), _p(levels["offsets"]), _p(levels["sizes"]),
                               _p(levels["ress"]), _p(levels["scales"]), _p(out))
    return out
