"""tests/golden/make_golden_128.py -- the BENCHMARK-SIZE render fixture (128x128 rays, max_steps 512, k = 13 evaluations per
sample, lambertian shading) computed ONCE by the CPU oracle (oracle/field_ref.py::render_train_ref, itself pinned against the
reference's Python by the 24x24 fixtures of make_golden.py) and committed as tests/golden/render_128.npz.
Forward quantities only (image, depth, weights_sum, the two regulariser means, per-ray sample counts): ~7 M field evaluations
take a few minutes on 8 cores; the backward at this size is covered by size-independent properties.
    python tests/golden/make_golden_128.py"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import camera_rays, field_from_golden, sphere_bitfield   # noqa: E402
from oracle import field_ref as fr                                     # noqa: E402


def main():
    g = dict(np.load(os.path.join(HERE, "render_albedo.npz")))
    field, _ = field_from_golden(g)
    HW = 128
    ro, rd, sc = camera_rays(HW)                         # radius 1.25, theta 80, phi 170, fov 20 (tests/helpers.py)
    bits = sphere_bitfield(0.2)
    rng = np.random.default_rng(128)
    noises = rng.random(HW * HW, dtype=np.float32)
    light = np.array([0.3, 0.5, 0.81], np.float32); light /= np.linalg.norm(light)
    bg = np.array([0.2, 0.5, 0.7], np.float32)
    t0 = time.time()
    with torch.no_grad():
        # chunk the rays so the oracle's [m, 13, 32] intermediates stay in memory; the smooth noise is drawn per chunk and saved
        outs, smooth = [], []
        CH = 2048
        for c0 in range(0, HW * HW, CH):
            sl = slice(c0, c0 + CH)
            # first pass to know m (cheap: march only)
            probe = fr.orm.march_rays_train(ro[sl], rd[sl], 1.0, bits, 1, 128, *fr.orm.near_far_from_aabb(ro[sl], rd[sl], np.array([-1, -1, -1, 1, 1, 1], np.float32), 0.2),
                                            noises[sl], 0.0, 512, align=128)
            m = probe[0].shape[0]
            sn = rng.standard_normal((m, 3)).astype(np.float32)
            o = fr.render_train_ref(field, ro[sl], rd[sl], bits, noises=noises[sl], light_d=light, smooth_noise=sn, bg_color=bg, depth_scale=sc[sl],
                                    max_steps=512, shading="lambertian", ambient_ratio=0.1, lambda_smooth=1.0)
            outs.append(o); smooth.append(sn)
            print(f"rays {c0 + CH}/{HW * HW}  samples {o['total']}  {time.time() - t0:.0f}s", flush=True)
    cat = lambda k: torch.cat([o[k] for o in outs]).numpy()
    counts = np.concatenate([o["rays"].numpy()[:, 2] for o in outs])
    # per-chunk padded means -> sums, so that the test can rebuild the whole-image means with the whole image's padded count
    sum_orient = sum(float(o["loss_orient"]) * o["xyzs"].shape[0] for o in outs)
    np.savez_compressed(os.path.join(HERE, "render_128.npz"), image=cat("image"), depth=cat("depth"), weights_sum=cat("weights_sum"),
                        counts=counts.astype(np.int32), total=np.int64(counts.sum()), noises=noises, light_d=light, bg_color=bg,
                        sum_orient=np.float64(sum_orient), radius=np.float32(0.2), ratio=np.float32(0.1))
    print("wrote render_128.npz", time.time() - t0, "s; total samples", int(counts.sum()))


if __name__ == "__main__":
    main()
