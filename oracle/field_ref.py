"""oracle/field_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-PyTorch (CPU, fp32, autograd) restatement of the field the reference evaluates per sample:

* ``HashGridRef``     tiny-cuda-nn HashGrid encoder as instantiated at nerf/network_tcnn.py:54-65.
                      tiny-cuda-nn is a THIRD-PARTY dependency that is NOT in /root/reference and is
                      un-pinned (README.md:43, git HEAD).  Restated from its published algorithm
                      (include/tiny-cuda-nn/encodings/grid.h, common_device.h).  PARITY UNPINNED:
                      no tcnn install and no golden vectors exist for it; it is cross-checked only
                      against the independent C restatement in raymarch_oracle.c.
* ``MLPRef``          nerf/network_tcnn.py:13-32
* ``FieldRef``        nerf/network_tcnn.py:94-170 (gaussian blob, common_forward, finite-difference
                      normal, shading) + activation.py:5-18 (trunc_exp) + nerf/utils.py:47-48
* ``render_train_ref``nerf/renderer.py:481-524,553-583 (training branch of run_cuda) on top of the
                      C oracle's march/composite (raymarching/src/raymarching.cu)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import raymarch as orm

PRIME_Y, PRIME_Z = 2654435761, 805459861
U32 = 0xFFFFFFFF


class _TruncExp(torch.autograd.Function):
    """activation.py:5-16 : fwd exp(x); bwd g * exp(clamp(x, max=15))."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(max=15))


def safe_normalize(x, eps=1e-20):
    """nerf/utils.py:47-48"""
    return x / torch.sqrt(torch.clamp(torch.sum(x * x, -1, keepdim=True), min=eps, max=1e32))


class HashGridRef(nn.Module):
    def __init__(self, n_levels=16, n_features=2, log2_hashmap_size=19, base_resolution=16, per_level_scale=None,
                 bound=1.0, seed=None, init_scale=1e-4):
        super().__init__()
        assert n_features == 2
        if per_level_scale is None:
            per_level_scale = float(np.exp2(np.log2(2048 * bound / 16) / (16 - 1)))  # network_tcnn.py:52
        self.levels = orm.hashgrid_levels(n_levels, base_resolution, per_level_scale, log2_hashmap_size)
        self.n_levels = n_levels
        n = self.levels["total"] * 2
        g = torch.Generator().manual_seed(0 if seed is None else seed)
        # tcnn initialises the grid U(-1e-4, 1e-4)
        self.params = nn.Parameter((torch.rand(n, generator=g) * 2 - 1) * init_scale)

    @property
    def n_output_dims(self):
        return 2 * self.n_levels

    def forward(self, x):
        """x [E,3] in [0,1] -> [E, 2L]; differentiable w.r.t. self.params only (tcnn gives no dx here)."""
        x = x.detach()
        E = x.shape[0]
        table = self.params.view(-1, 2)
        outs = []
        for l in range(self.n_levels):
            scale = float(self.levels["scales"][l])
            res = int(self.levels["ress"][l])
            size = int(self.levels["sizes"][l])
            off = int(self.levels["offsets"][l])
            pos = (x.double() * scale + 0.5).float()          # fmaf(scale, x, 0.5f) emulated via one fp64 op
            fl = torch.floor(pos)
            w = pos - fl                                      # [E,3]
            cell = fl.to(torch.int64)
            feat = torch.zeros(E, 2, dtype=torch.float32)
            for corner in range(8):
                wt = torch.ones(E, dtype=torch.float32)
                c = []
                for d in range(3):
                    if corner & (1 << d):
                        wt = wt * w[:, d]
                        c.append((cell[:, d] + 1) & U32)
                    else:
                        wt = wt * (1 - w[:, d])
                        c.append(cell[:, d] & U32)
                # grid_index(): dense walk while stride <= size, hashed if size < final stride
                stride, idx = 1, torch.zeros(E, dtype=torch.int64)
                for d in range(3):
                    if stride > size:
                        break
                    idx = (idx + c[d] * stride) & U32
                    stride *= res
                if size < stride:
                    idx = c[0] ^ ((c[1] * PRIME_Y) & U32) ^ ((c[2] * PRIME_Z) & U32)
                idx = idx % size
                feat = feat + wt[:, None] * table[off + idx]
            outs.append(feat)
        return torch.cat(outs, dim=-1)


class MLPRef(nn.Module):
    """nerf/network_tcnn.py:13-32 (same parameter names: net.{l}.weight / net.{l}.bias)."""

    def __init__(self, dim_in=32, dim_out=4, dim_hidden=64, num_layers=3):
        super().__init__()
        self.num_layers = num_layers
        self.net = nn.ModuleList([
            nn.Linear(dim_in if l == 0 else dim_hidden, dim_out if l == num_layers - 1 else dim_hidden, bias=True)
            for l in range(num_layers)])

    def forward(self, x):
        for l in range(self.num_layers):
            x = self.net[l](x)
            if l != self.num_layers - 1:
                x = F.relu(x)
        return x


class FieldRef(nn.Module):
    """nerf/network_tcnn.py:37-205 minus the renderer base class."""

    def __init__(self, bound=1.0, blob_density=5.0, blob_radius=0.1, seed=0, table_scale=1e-4):
        super().__init__()
        self.bound = bound
        self.blob_density, self.blob_radius = blob_density, blob_radius
        self.encoder = HashGridRef(bound=bound, seed=seed, init_scale=table_scale)
        torch.manual_seed(seed)
        self.sigma_net = MLPRef(32, 4, 64, 3)

    def gaussian(self, x):                                  # :94-100
        d = (x ** 2).sum(-1)
        return self.blob_density * torch.exp(-d / (2 * self.blob_radius ** 2))

    def common_forward(self, x):                            # :102-112
        h = (x + self.bound) / (2 * self.bound)
        h = self.encoder(h)
        h = self.sigma_net(h)
        sigma = _TruncExp.apply(h[..., 0] + self.gaussian(x))
        albedo = torch.sigmoid(h[..., 1:])
        return sigma, albedo

    def finite_difference_normal(self, x, epsilon=1e-2):    # :115-130
        taps = []
        for axis in range(3):
            e = torch.zeros(1, 3)
            e[0, axis] = epsilon
            pos, _ = self.common_forward((x + e).clamp(-self.bound, self.bound))
            neg, _ = self.common_forward((x - e).clamp(-self.bound, self.bound))
            taps.append(0.5 * (pos - neg) / epsilon)
        return -torch.stack(taps, dim=-1)

    def normal(self, x):                                    # :132-138
        n = self.finite_difference_normal(x)
        n = safe_normalize(n)
        return torch.nan_to_num(n)

    def forward(self, x, d, l=None, ratio=1, shading='albedo'):   # :140-170
        sigma, albedo = self.common_forward(x)
        normal = self.normal(x)
        if shading == 'albedo':
            color = albedo
        elif normal.shape[0] < 1e6:
            lambertian = ratio + (1 - ratio) * (normal @ l).clamp(min=0.1)
            if shading == 'textureless':
                color = lambertian.unsqueeze(-1).repeat(1, 3)
            elif shading == 'normal':
                color = (normal + 1) / 2
            else:
                color = albedo * lambertian.unsqueeze(-1)
        else:
            color = albedo
        return sigma, color, normal

    def density(self, x):                                   # :173-180
        sigma, albedo = self.common_forward(x)
        return {'sigma': sigma, 'albedo': albedo}


class _CompositeTrainRef(torch.autograd.Function):
    """raymarching/raymarching.py:250-300 over the C oracle."""

    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh):
        ws, depth, image = orm.composite_rays_train_forward(sigmas.detach().numpy(), rgbs.detach().numpy(),
                                                            deltas.numpy(), rays.numpy(), T_thresh)
        ws, depth, image = torch.from_numpy(ws), torch.from_numpy(depth), torch.from_numpy(image)
        ctx.save_for_backward(sigmas.detach(), rgbs.detach(), deltas, rays, ws, image)
        ctx.T_thresh = T_thresh
        return ws, depth, image

    @staticmethod
    def backward(ctx, g_ws, g_depth, g_image):
        sigmas, rgbs, deltas, rays, ws, image = ctx.saved_tensors
        gs, gr = orm.composite_rays_train_backward(g_ws.contiguous().numpy(), g_image.contiguous().numpy(),
                                                   sigmas.numpy(), rgbs.numpy(), deltas.numpy(), rays.numpy(),
                                                   ws.numpy(), image.numpy(), ctx.T_thresh)
        return torch.from_numpy(gs), torch.from_numpy(gr), None, None, None


def render_train_ref(field, rays_o, rays_d, bitfield, *, cascade=1, grid_size=128, bound=1.0, noises, light_d,
                     smooth_noise=None, bg_color=None, depth_scale=None, dt_gamma=0.0, max_steps=512, T_thresh=1e-4,
                     ambient_ratio=1.0, shading='albedo', lambda_smooth=1.0, min_near=0.2, max_depth=10.0, aabb=None):
    """Training branch of NeRFRenderer.run_cuda (nerf/renderer.py:481-524, 553-583) with every random draw injected:
    noises[N] (raymarching.py:226), light_d[3] (renderer.py:498), smooth_noise[m,3] (renderer.py:522, already N(0,1),
    scaled by 1e-2 here), bg_color[3] (utils.py:491).  Returns a dict like run_cuda plus the marched samples."""
    rays_o = torch.as_tensor(rays_o, dtype=torch.float32).reshape(-1, 3)
    rays_d = torch.as_tensor(rays_d, dtype=torch.float32).reshape(-1, 3)
    N = rays_o.shape[0]
    if aabb is None:
        aabb = np.array([-bound, -bound, -bound, bound, bound, bound], np.float32)
    nears, fars = orm.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb, min_near)
    xyzs, dirs, deltas, rays, total = orm.march_rays_train(rays_o.numpy(), rays_d.numpy(), bound, bitfield, cascade,
                                                           grid_size, nears, fars, noises, dt_gamma, max_steps, align=128)
    xyzs, dirs, deltas, rays = map(torch.from_numpy, (xyzs, dirs, deltas, rays))
    light_d = torch.as_tensor(light_d, dtype=torch.float32)
    sigmas, rgbs, normals = field(xyzs, dirs, light_d, ratio=ambient_ratio, shading=shading)
    ws, depth, image = _CompositeTrainRef.apply(sigmas, rgbs, deltas, rays, T_thresh)
    out = {}
    weights = 1 - torch.exp(-sigmas)
    out['loss_orient'] = (weights.detach() * (normals * dirs).sum(-1).clamp(min=0) ** 2).mean()
    if lambda_smooth > 0:
        if smooth_noise is None:
            smooth_noise = torch.randn_like(xyzs)
        smooth_noise = torch.as_tensor(smooth_noise, dtype=torch.float32)
        normals_perturb = field.normal(xyzs + smooth_noise * 1e-2)
        out['loss_smooth'] = (normals - normals_perturb).abs().mean()
    bg = 1 if bg_color is None else torch.as_tensor(bg_color, dtype=torch.float32)
    image = image + (1 - ws).unsqueeze(-1) * bg
    depth = depth + (1 - ws) * max_depth
    if depth_scale is not None:
        depth = depth * torch.as_tensor(depth_scale, dtype=torch.float32).reshape(-1)
    out.update(image=image, depth=depth, weights_sum=ws, mask=torch.from_numpy(nears < fars),
               xyzs=xyzs, dirs=dirs, deltas=deltas, rays=rays, total=total, sigmas=sigmas, rgbs=rgbs, normals=normals,
               nears=torch.from_numpy(nears), fars=torch.from_numpy(fars))
    return out


def get_rays_ref(pose, intrinsics, H, W):
    """nerf/utils.py:51-116 for N=-1 (all pixels), B=1.  pose [4,4] cam2world, intrinsics (fx,fy,cx,cy)."""
    pose = torch.as_tensor(pose, dtype=torch.float32)
    fx, fy, cx, cy = intrinsics
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing='ij')
    i = i.t().reshape(H * W) + 0.5
    j = j.t().reshape(H * W) + 0.5
    zs = torch.ones_like(i)
    dirs = torch.stack(((i - cx) / fx * zs, (j - cy) / fy * zs, zs), dim=-1)
    scale = 1 / dirs.pow(2).sum(-1).pow(0.5)
    dirs = safe_normalize(dirs)
    rays_d = dirs @ pose[:3, :3].t()
    rays_o = pose[:3, 3].expand_as(rays_d)
    return rays_o.contiguous(), rays_d.contiguous(), scale


def orbit_pose(radius, theta_deg, phi_deg):
    """Camera on a sphere looking at the origin, the convention of nerf/provider.py:143-214 (fix_poses):
    centers = r*(sin(th)sin(ph), cos(th), sin(th)cos(ph)); forward = -normalize(centers) ... (look-at, up=+y)."""
    th, ph = math.radians(theta_deg), math.radians(phi_deg)
    c = np.array([radius * math.sin(th) * math.sin(ph), radius * math.cos(th), radius * math.sin(th) * math.cos(ph)], np.float32)
    fwd = -c / np.linalg.norm(c)
    up = np.array([0, -1, 0], np.float32)                  # provider.py:203
    right = np.cross(fwd, up); right /= np.linalg.norm(right)
    up2 = np.cross(right, fwd); up2 /= np.linalg.norm(up2)
    pose = np.eye(4, dtype=np.float32)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = right, up2, fwd, c
    return pose
