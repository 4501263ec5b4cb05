"""B2 seam: NeRFRenderer with the reference's attribute / buffer / method names (nerf/renderer.py:99-677).

Only the CUDA-ray path exists (main.py:95 hard-sets cuda_ray=True; the reference's pure-PyTorch `run()` is dead code).
`run_cuda`'s training branch is three launches of libmi3d.so with no host synchronisation; `update_extra_state`
keeps the mean density and the packing threshold on the device.  state_dict keys (density_grid, density_bitfield,
step_counter, aabb_train, aabb_infer) match the reference so checkpoints interchange (nerf/utils.py:1075-1186).
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from .. import _lib as L
from .. import raymarching
from . import field_ops
from .utils import safe_normalize


class NeRFRenderer(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.bound = opt.bound
        self.cascade = 1 + math.ceil(math.log2(opt.bound))
        self.grid_size = 128
        self.cuda_ray = True
        self.min_near = opt.min_near
        self.density_thresh = opt.density_thresh
        self.bg_radius = getattr(opt, "bg_radius", -1)
        aabb_train = torch.FloatTensor([-opt.bound, -opt.bound, -opt.bound, opt.bound, opt.bound, opt.bound])
        self.register_buffer('aabb_train', aabb_train)
        self.register_buffer('aabb_infer', aabb_train.clone())
        self.register_buffer('density_grid', torch.zeros([self.cascade, self.grid_size ** 3]))
        self.register_buffer('density_bitfield', torch.zeros(self.cascade * self.grid_size ** 3 // 8, dtype=torch.uint8))
        self.register_buffer('step_counter', torch.zeros(16, 2, dtype=torch.int32))
        self.mean_density = 0
        self.iter_density = 0
        self.mean_count = 0
        self.local_step = 0
        self._mean_density_dev = None
        self._workspaces = {}
        self._grid_ws = None

    # --- to be provided by the field subclass ---
    def forward(self, x, d):
        raise NotImplementedError()

    def density(self, x):
        raise NotImplementedError()

    def _field_handles(self):
        """-> (table, (w1,b1,w2,b2,w3,b3), hashgrid struct, base cfg dict)"""
        raise NotImplementedError()

    def reset_extra_state(self):
        self.density_grid.zero_()
        self.mean_density = 0
        self.iter_density = 0
        self.step_counter.zero_()
        self.mean_count = 0
        self.local_step = 0

    def _workspace(self, N, max_steps, device):
        key = (N, max_steps, str(device))
        ws = self._workspaces.get(key)
        if ws is None:
            ws = field_ops.RenderWorkspace(N, max_steps, device, max_samples=getattr(self.opt, "max_samples", None))
            self._workspaces = {key: ws}        # keep one (shapes are fixed during training)
        return ws

    def run_cuda(self, rays_o, rays_d, depth_scale=None, bg_color=None, dt_gamma=0, light_d=None, ambient_ratio=1.0,
                 shading='albedo', perturb=False, force_all_rays=False, max_steps=1024, T_thresh=1e-4, **kwargs):
        """nerf/renderer.py:481-583.  Extra optional kwargs (additive): noises[N], smooth_noise[m,3] inject the random draws
        of raymarching.py:226 / renderer.py:522 for reproducible parity tests."""
        prefix = rays_o.shape[:-1]
        rays_o = L.f32c(rays_o).view(-1, 3)
        rays_d = L.f32c(rays_d).view(-1, 3)
        L.require_cuda(rays_o, rays_d)
        N = rays_o.shape[0]
        device = rays_o.device

        if light_d is None:                                            # renderer.py:496-499
            light_d = safe_normalize(rays_o[0] + torch.randn(3, device=device, dtype=torch.float))
        light_d = L.f32c(light_d)
        results = {}

        if self.training:
            table, mlp_params, hg, cfg = self._field_handles()
            cfg = dict(cfg, n_evals=13 if self.opt.lambda_smooth > 0 else 7, shading=shading, ambient_ratio=float(ambient_ratio))
            ws = self._workspace(N, max_steps, device)
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())          # CPU generator: no device sync
            opts = dict(bound=float(self.bound), dt_gamma=float(dt_gamma), max_steps=int(max_steps), cascade=self.cascade,
                        grid_size=self.grid_size, min_near=0.2,          # wrapper default, NOT opt.min_near (raymarching.py:34)
                        T_thresh=float(T_thresh), max_depth=float(self.opt.max_depth), seed=seed)
            noises = kwargs.get('noises')
            if noises is None and not perturb:
                noises = torch.zeros(N, dtype=torch.float32, device=device)
            if bg_color is not None and not torch.is_tensor(bg_color):
                bg_color = torch.full((3,), float(bg_color), device=device)
            if bg_color is not None:
                bg_color = L.f32c(bg_color.to(device))
            ds = L.f32c(depth_scale).view(-1) if depth_scale is not None else None
            self.local_step += 1
            image, depth, weights_sum, loss_orient, loss_smooth = field_ops.render_train(
                table, mlp_params, rays_o, rays_d, self.density_bitfield, self.aabb_train, ws, hg, cfg, opts,
                noises=noises, light_d=light_d, smooth_noise=kwargs.get('smooth_noise'), bg_color=bg_color, depth_scale=ds)
            results['loss_orient'] = loss_orient
            if self.opt.lambda_smooth > 0:
                results['loss_smooth'] = loss_smooth
            nears, fars = ws.nears, ws.fars
            image = image.view(*prefix, 3)
            depth = depth.view(*prefix, 1)
        else:
            nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, self.aabb_infer)
            dtype = torch.float32
            weights_sum = torch.zeros(N, dtype=dtype, device=device)
            depth = torch.zeros(N, dtype=dtype, device=device)
            image = torch.zeros(N, 3, dtype=dtype, device=device)
            normal = torch.zeros(N, 3, dtype=dtype, device=device)
            n_alive = N
            rays_alive = torch.arange(n_alive, dtype=torch.int32, device=device)
            rays_t = nears.clone()
            step = 0
            while step < max_steps:                                   # renderer.py:535-551
                n_alive = rays_alive.shape[0]
                if n_alive <= 0:
                    break
                n_step = max(min(N // n_alive, 8), 1)
                xyzs, dirs, deltas = raymarching.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.bound,
                                                            self.density_bitfield, self.cascade, self.grid_size, nears, fars, 128,
                                                            perturb if step == 0 else False, dt_gamma, max_steps)
                sigmas, rgbs, normals = self(xyzs, dirs, light_d, ratio=ambient_ratio, shading=shading)
                normals = (normals + 1) / 2
                raymarching.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, normals, deltas, weights_sum, depth,
                                           image, normal, T_thresh)
                rays_alive = rays_alive[rays_alive >= 0]
                step += n_step
            bg = 1 if bg_color is None else bg_color
            image = (image + (1 - weights_sum).unsqueeze(-1) * bg).view(*prefix, 3)
            normal = (normal + (1 - weights_sum).unsqueeze(-1) * bg).view(*prefix, 3)
            depth = depth + (1 - weights_sum) * self.opt.max_depth
            if depth_scale is not None:
                depth = depth.view(*prefix, 1) * depth_scale.view(*prefix, 1)
            else:
                depth = depth.view(*prefix, 1)
            results['normal'] = normal

        results['image'] = image
        results['depth'] = depth
        results['weights_sum'] = weights_sum.reshape(*prefix)
        results['mask'] = (nears < fars).reshape(*prefix)
        return results

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128, jitter=None):
        """nerf/renderer.py:587-637 as ONE C-ABI call (positions, field, EMA-max, mean, packbits); no .item()."""
        table, mlp_params, hg, cfg = self._field_handles()
        dev = table.device
        lib = L.lib()
        if self._grid_ws is None or self._grid_ws.device != dev:
            nbytes = lib.mi3d_density_grid_workspace_bytes(C.c_uint32(self.cascade), C.c_uint32(self.grid_size))
            self._grid_ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._mean_density_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        cf = field_ops._cfg_struct(dict(cfg, n_evals=1, shading='albedo', ambient_ratio=1.0), None)
        mlp = field_ops._mlp_struct(mlp_params)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        if jitter is not None:
            jitter = L.f32c(jitter)
        L.check(lib.mi3d_density_grid_update(
            L.ptr(self.density_grid), L.ptr(self.density_bitfield), C.c_uint32(self.cascade), C.c_uint32(self.grid_size),
            C.c_float(self.bound), C.c_float(decay), C.c_float(self.density_thresh), L.ptr(table), C.byref(hg), C.byref(mlp),
            C.byref(cf), L.ptr(jitter), C.c_uint64(seed), L.ptr(self._mean_density_dev), L.ptr(self._grid_ws), L.stream()),
            "density_grid_update")
        self.mean_density = self._mean_density_dev     # device scalar; float(...) only when a checkpoint is written
        self.iter_density += 1
        self.mean_count = 0                            # unused under force_all_rays=True (nerf/utils.py:498)
        self.local_step = 0

    def render(self, rays_o, rays_d, depth_scale=None, staged=False, max_ray_batch=4096, **kwargs):
        """nerf/renderer.py:642-677 (cuda_ray never stages)."""
        return self.run_cuda(rays_o, rays_d, depth_scale, **kwargs)
