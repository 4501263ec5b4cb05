"""tests/golden/make_golden_r6.py -- fixture for row R6 (density-grid refresh) from the REFERENCE's own
NeRFRenderer.update_extra_state (nerf/renderer.py:587-637), run on the CPU through the import shims of make_golden.py
(raymarching -> C oracle, tinycudann -> oracle HashGridRef; everything else is the reference's code).
The cell jitter is the reference's own `torch.rand_like` draw: seeded with torch.manual_seed(SEED) right before the call, so the
test regenerates it with torch.manual_seed(SEED); torch.rand(128**3, 3) (CPU generator) instead of storing 25 MB.
Stored: the initial grid seed, 40 000 sampled cells of the updated grid, the mean density, the packed bitfield.
    python tests/golden/make_golden_r6.py"""
import argparse
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg          # noqa: E402  (shims + helpers)

SEED = 4321


def main():
    mg.install_shims()
    rm = sys.modules["raymarching"]
    orm = mg.orm
    rm.morton3D = lambda coords: torch.from_numpy(orm.morton3D(coords.numpy().astype(np.int32)))
    rm.packbits = lambda grid, thresh, bitfield=None: torch.from_numpy(orm.packbits(grid.numpy().reshape(-1), float(thresh)))
    from nerf.network_tcnn import NeRFNetwork           # REFERENCE class
    g = dict(np.load(os.path.join(HERE, "render_albedo.npz")))
    opt = argparse.Namespace(bound=1, min_near=0.1, density_thresh=10, bg_radius=-1, blob_density=5, blob_radius=0.1, cuda_ray=True,
                             lambda_smooth=1, max_depth=10.0)
    torch.manual_seed(0)
    net = NeRFNetwork(opt)
    table = mg.make_table(net.encoder.params.numel(), int(g["table_seed"]), float(g["table_scale"]))
    with torch.no_grad():
        net.encoder.params.copy_(torch.from_numpy(table))
        for l, (w, b) in enumerate((("w1", "b1"), ("w2", "b2"), ("w3", "b3"))):
            net.sigma_net.net[l].weight.copy_(torch.from_numpy(g[w]))
            net.sigma_net.net[l].bias.copy_(torch.from_numpy(g[b]))
    H = 128
    grid0 = (np.random.default_rng(21).random((1, H ** 3), dtype=np.float32) * 2).astype(np.float32)
    grid0[0, ::97] = -1.0                                   # a few cells marked invalid (< 0): the EMA must leave them alone
    net.density_grid.copy_(torch.from_numpy(grid0))
    t0 = time.time()
    torch.manual_seed(SEED)
    net.update_extra_state(decay=0.95)
    print(f"reference update_extra_state on CPU: {time.time() - t0:.1f} s; mean_density {net.mean_density}")
    grid = net.density_grid.numpy()[0]
    sub = np.random.default_rng(5).choice(H ** 3, 40000, replace=False)
    np.savez_compressed(os.path.join(HERE, "density_r6.npz"), seed=np.int64(SEED), grid0_seed=np.int64(21), sub=sub.astype(np.int32),
                        sub_vals=grid[sub], mean_density=np.float32(net.mean_density), bitfield=net.density_bitfield.numpy().astype(np.uint8))
    print("wrote density_r6.npz")


if __name__ == "__main__":
    main()
