"""CUDA-event timing of the SD guidance path at full SD-2.0-base size (seeded random weights): encode, U-Net+CFG+SDS, encode backward."""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    sdm = importlib.import_module("make-it-3d_b200.nerf.sd")
    t0 = time.time()
    g = sdm.StableDiffusion("cuda")
    torch.cuda.synchronize()
    print(f"engine built in {time.time() - t0:.1f}s, workspace {g.engine.nbytes / 2**30:.2f} GiB, params {len(g.engine.names)}")
    rgb = torch.rand(1, 3, 128, 128, device="cuda", requires_grad=True)
    ctx = torch.randn(2, 77, 1024, device="cuda")
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    te, tu, tb = [], [], []
    side = torch.cuda.Stream() if os.environ.get("PROF_SIDE_STREAM") else None     # a capturable stream (MI3D_SD_GRAPH=1 experiments)
    if side is not None:
        side.wait_stream(torch.cuda.current_stream())
        torch.cuda.set_stream(side)
    for it in range(iters + 2):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        rgb.grad = None
        g._t.fill_(500)
        e[0].record()
        lat = g.encode_imgs(rgb)
        e[1].record()
        noise = torch.randn_like(lat)
        npred, grad = g.unet_sds(lat, noise, g._t, ctx, 10.0)
        e[2].record()
        lat.backward(gradient=grad)
        e[3].record()
        torch.cuda.synchronize()
        if it >= 2:
            te.append(e[0].elapsed_time(e[1])); tu.append(e[1].elapsed_time(e[2])); tb.append(e[2].elapsed_time(e[3]))
    med = lambda v: float(np.median(v))
    print(f"VAE encode fwd {med(te):.3f} ms | U-Net(x2)+CFG+SDS {med(tu):.3f} ms | VAE encode bwd {med(tb):.3f} ms | finite: {bool(torch.isfinite(rgb.grad).all())} {bool(torch.isfinite(npred).all())}")
    if os.environ.get("MI3D_GEMM_DUMP"):
        gemm_dump(g, rgb, ctx)
    print(f"tensor-roofline: unet {1.608e12 / (med(tu) * 1e-3) / 1e12:.1f} TFLOP/s, vae fwd {1.117e12 / (med(te) * 1e-3) / 1e12:.1f}, vae bwd {1.117e12 / (med(tb) * 1e-3) / 1e12:.1f}")


def gemm_dump(g, rgb, ctx):
    """Per-launch list of the tile kernel (CUDA events around each launch) for one guidance step -> gpurun_out/sd_gemms.txt"""
    import ctypes as C
    L = importlib.import_module("make-it-3d_b200._lib")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    L.check(L.lib().mi3d_sd_profile(g.engine.h, C.c_int(1), None, None), "sd_profile")
    rgb.grad = None
    lat = g.encode_imgs(rgb)
    npred, grad = g.unet_sds(lat, torch.randn_like(lat), g._t, ctx, 10.0)
    lat.backward(gradient=grad)
    torch.cuda.synchronize()
    ms, n = C.c_float(0), C.c_int(0)
    L.check(L.lib().mi3d_sd_profile_dump(g.engine.h, C.c_int(0), C.byref(ms), C.byref(n), os.path.join(ROOT, "gpurun_out", "sd_gemms.txt").encode()), "sd_profile_dump")
    print(f"tile kernel: {n.value} launches, {ms.value:.3f} ms")


if __name__ == "__main__":
    main()
