"""tests/golden/make_golden.py -- generates the committed golden fixtures by running the REFERENCE's own Python.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

What is real reference code here (imported unmodified from /root/reference):
    nerf/network_tcnn.py  NeRFNetwork.{common_forward, finite_difference_normal, normal, forward, gaussian}, MLP
    nerf/renderer.py      NeRFRenderer.run_cuda (training branch: regularisers, bg mix, depth fix-up), render
    activation.py         trunc_exp
    nerf/utils.py         safe_normalize, get_rays
What is substituted, because it cannot run here (no GPU, third-party packages absent):
    `raymarching` module  -> shim over the C oracle (oracle/raymarch_oracle.c); the reference wrappers call .cuda()
    `tinycudann`          -> stub whose Encoding is oracle.field_ref.HashGridRef (tcnn is un-pinned third party)
    11 import-only deps   -> MagicMock
Random draws are injected/recorded: march noises, light_d, bg_color, smooth-loss noise (torch.manual_seed before render).
"""
import argparse
import math
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import field_ref as fr          # noqa: E402
from oracle import raymarch as orm          # noqa: E402

STATE = {}


def install_shims():
    for name in ["trimesh", "open3d", "mcubes", "imageio", "tensorboardX", "matplotlib", "matplotlib.pyplot", "torch_ema", "clip",
                 "torchmetrics", "contextual_loss", "pytorch3d", "pytorch3d.renderer", "pytorch3d.renderer.compositing",
                 "pytorch3d.renderer.points", "pytorch3d.renderer.points.rasterize_points", "nvdiffrast", "nvdiffrast.torch",
                 "xatlas", "cv2", "rich", "rich.console", "tqdm", "lpips", "pytorch3d.structures", "pytorch3d.ops",
                 "sklearn", "sklearn.neighbors", "scipy.ndimage", "skimage", "PIL", "PIL.Image", "torchvision", "torchvision.utils",
                 "torchvision.transforms", "pandas", "kornia", "trimesh.exchange", "pytorch3d.renderer.points.rasterizer"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock()
    # tinycudann stub
    tcnn = types.ModuleType("tinycudann")

    class Encoding(torch.nn.Module):
        def __init__(self, n_input_dims, encoding_config, dtype=torch.float32):
            super().__init__()
            self.inner = fr.HashGridRef(n_levels=encoding_config["n_levels"], log2_hashmap_size=encoding_config["log2_hashmap_size"],
                                        base_resolution=encoding_config["base_resolution"],
                                        per_level_scale=float(encoding_config["per_level_scale"]))
            self.params = self.inner.params

        def forward(self, x):
            return self.inner(x)

    tcnn.Encoding = Encoding
    sys.modules["tinycudann"] = tcnn

    # raymarching shim over the C oracle (CPU tensors)
    rm = types.ModuleType("raymarching")

    def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
        n, f = orm.near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb.numpy(), min_near)
        return torch.from_numpy(n), torch.from_numpy(f)

    def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                         perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        noises = STATE["noises"] if perturb else np.zeros(rays_o.shape[0], np.float32)
        xyzs, dirs, deltas, rays, total = orm.march_rays_train(rays_o.numpy(), rays_d.numpy(), bound, density_bitfield.numpy(), C, H,
                                                               nears.numpy(), fars.numpy(), noises, dt_gamma, max_steps, align=align)
        STATE["total"] = total
        STATE["m_pad"] = xyzs.shape[0]
        return tuple(map(torch.from_numpy, (xyzs, dirs, deltas, rays)))

    def composite_rays_train(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        return fr._CompositeTrainRef.apply(sigmas, rgbs, deltas, rays, T_thresh)

    rm.near_far_from_aabb = near_far_from_aabb
    rm.march_rays_train = march_rays_train
    rm.composite_rays_train = composite_rays_train
    sys.modules["raymarching"] = rm
    sys.path.insert(0, REF)


def sphere_bitfield(radius, H=128):
    idx = np.arange(H ** 3, dtype=np.int32)
    coords = orm.morton3D_invert(idx)
    xyz = (coords.astype(np.float32) + 0.5) / H * 2 - 1
    grid = (np.linalg.norm(xyz, axis=1) < radius).astype(np.float32)
    return orm.packbits(grid, 0.5)


def make_table(n_params, seed, scale):
    return ((np.random.default_rng(seed).random(n_params, dtype=np.float32) * 2 - 1) * scale).astype(np.float32)


def make_case(name, shading, ratio, HW=24, radius=0.3, seed=0, table_scale=0.5):
    from nerf.network_tcnn import NeRFNetwork           # REFERENCE class
    opt = argparse.Namespace(bound=1, min_near=0.1, density_thresh=10, bg_radius=-1, blob_density=5, blob_radius=0.1, cuda_ray=True,
                             lambda_smooth=1, max_depth=10.0)
    torch.manual_seed(seed)
    net = NeRFNetwork(opt)
    net.train()
    table = make_table(net.encoder.params.numel(), 1234 + seed, table_scale)
    with torch.no_grad():
        net.encoder.params.copy_(torch.from_numpy(table))
        # push the output layer away from init so densities / colours vary across the object
        net.sigma_net.net[2].weight.mul_(3.0)
        net.sigma_net.net[2].bias[0] = 1.5
    bits = sphere_bitfield(radius)
    net.density_bitfield = torch.from_numpy(bits)

    rng = np.random.default_rng(77 + seed)
    pose = fr.orbit_pose(1.2, 78.0, 160.0 + 20 * seed)
    focal = HW / (2 * math.tan(math.radians(20.0) / 2))
    from nerf.utils import get_rays                       # REFERENCE function
    rays = get_rays(torch.from_numpy(pose)[None], (focal, focal, HW / 2, HW / 2), HW, HW, -1)
    rays_o, rays_d, depth_scale = rays['rays_o'].contiguous(), rays['rays_d'].contiguous(), rays['depth_scale'].contiguous()
    N = HW * HW
    STATE["noises"] = rng.random(N, dtype=np.float32)
    light_d = fr.safe_normalize(torch.from_numpy(rng.standard_normal(3).astype(np.float32)) + rays_o[0, 0])
    bg_color = torch.from_numpy(rng.random(3, dtype=np.float32))
    noise_seed = 4242 + seed
    torch.manual_seed(noise_seed)
    out = net.render(rays_o, rays_d, depth_scale=depth_scale, bg_color=bg_color, staged=False, perturb=True, light_d=light_d,
                     ambient_ratio=ratio, shading=shading, force_all_rays=True, max_steps=512, dt_gamma=0, T_thresh=1e-4)
    m_pad = STATE["m_pad"]
    torch.manual_seed(noise_seed)
    smooth_noise = torch.randn(m_pad, 3)                  # == the randn_like(xyzs) drawn at renderer.py:522

    A = torch.from_numpy(rng.standard_normal((N, 3)).astype(np.float32))
    B = torch.from_numpy(rng.standard_normal(N).astype(np.float32))
    Cd = torch.from_numpy(rng.standard_normal(N).astype(np.float32)) * 0.1
    loss = (out['image'][0] * A).sum() + (out['weights_sum'][0] * B).sum() + (out['depth'][0, :, 0] * Cd).sum() \
        + 30.0 * out['loss_orient'] + 50.0 * out['loss_smooth']
    net.zero_grad()
    loss.backward()
    gtab = net.encoder.params.grad.numpy()
    nz = np.flatnonzero(gtab)
    pick = np.sort(rng.choice(nz, size=min(4096, nz.size), replace=False))
    mlp = [p.detach().numpy() for p in net.sigma_net.parameters()]
    gmlp = [p.grad.numpy() for p in net.sigma_net.parameters()]
    np.savez_compressed(
        os.path.join(HERE, f"render_{name}.npz"),
        shading=shading, ratio=np.float32(ratio), HW=HW, radius=np.float32(radius), table_seed=1234 + seed,
        table_scale=np.float32(table_scale), rays_o=rays_o[0].numpy(), rays_d=rays_d[0].numpy(), depth_scale=depth_scale[0].numpy(),
        noises=STATE["noises"], light_d=light_d.numpy(), bg_color=bg_color.numpy(), smooth_noise=smooth_noise.numpy().astype(np.float32),
        total=STATE["total"], m_pad=m_pad,
        w1=mlp[0], b1=mlp[1], w2=mlp[2], b2=mlp[3], w3=mlp[4], b3=mlp[5],
        image=out['image'][0].detach().numpy(), depth=out['depth'][0, :, 0].detach().numpy(),
        weights_sum=out['weights_sum'][0].detach().numpy(), mask=out['mask'][0].numpy(),
        loss_orient=out['loss_orient'].item(), loss_smooth=out['loss_smooth'].item(),
        A=A.numpy(), B=B.numpy(), Cd=Cd.numpy(),
        g_w1=gmlp[0], g_b1=gmlp[1], g_w2=gmlp[2], g_b2=gmlp[3], g_w3=gmlp[4], g_b3=gmlp[5],
        g_table_idx=pick.astype(np.int64), g_table_val=gtab[pick], g_table_sum=np.float64(gtab.astype(np.float64).sum()),
        g_table_l2=np.float64(np.sqrt((gtab.astype(np.float64) ** 2).sum())), g_table_nnz=nz.size)
    print(name, "total", STATE["total"], "m_pad", m_pad, "loss_orient", out['loss_orient'].item(), "loss_smooth",
          out['loss_smooth'].item(), "ws mean", out['weights_sum'].mean().item(), "nnz", nz.size)


def make_small_ops():
    """trunc_exp / MLP / safe_normalize / get_rays vectors straight from the reference modules."""
    from activation import trunc_exp
    from nerf.utils import get_rays, safe_normalize
    x = torch.linspace(-20, 20, 41, requires_grad=True)
    y = trunc_exp(x)
    y.sum().backward()
    pose = torch.from_numpy(fr.orbit_pose(1.3, 85.0, 200.0))[None]
    r = get_rays(pose, (40.0, 42.0, 8.0, 7.5), 15, 16, -1)
    v = torch.tensor([[0.0, 0.0, 0.0], [1e-12, 0, 0], [3.0, 4.0, 0.0], [1e20, 1e20, 0]])
    np.savez_compressed(os.path.join(HERE, "small_ops.npz"), te_x=x.detach().numpy(), te_y=y.detach().numpy(), te_g=x.grad.numpy(),
                        pose=pose[0].numpy(), rays_o=r['rays_o'][0].numpy(), rays_d=r['rays_d'][0].numpy(),
                        depth_scale=r['depth_scale'][0].numpy(), sn_in=v.numpy(), sn_out=safe_normalize(v).numpy())


if __name__ == "__main__":
    install_shims()
    make_small_ops()
    make_case("albedo", "albedo", 1.0, seed=0)
    make_case("lambertian", "lambertian", 0.1, seed=1)
    make_case("textureless", "textureless", 0.1, seed=2)
