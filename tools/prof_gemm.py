"""Stand-alone launches of the tensor-core tile kernel / fused attention at SD-2.0-base shapes (for ncu captures and quick timing).
Usage (GPU box): python tools/prof_gemm.py [iters]"""
import ctypes as C
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [  # (M, N, K, what)
    (8192, 320, 320, "level-0 attention projection"),
    (2048, 640, 640, "level-1 attention projection"),
    (8192, 2560, 320, "level-0 GEGLU up-projection (plain epilogue here)"),
    (65536, 256, 2304, "VAE 256x256 conv as GEMM"),
]


def main():
    L = importlib.import_module("make-it-3d_b200._lib")
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    g = torch.Generator(device="cuda").manual_seed(0)
    for M, N, K, what in SHAPES:
        a = torch.randn(M, K, device="cuda", generator=g).half()
        b = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
        out = torch.empty(M, N, dtype=torch.float16, device="cuda")
        call = lambda: L.check(L.lib().mi3d_gemm_f16(L.ptr(a), L.ptr(b), L.ptr(out), C.c_int(0), C.c_int(M), C.c_int(N), C.c_int(K), C.c_int(0),
                                                     C.c_float(1.0), None, None, C.c_int(0), L.stream()), "gemm")
        for _ in range(3):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            call()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        print(f"gemm {M}x{N}x{K} ({what}): {us:.1f} us/launch back-to-back, {2.0 * M * N * K / us / 1e6:.1f} TFLOP/s")
    # fused attention, level 0: B=2, T=4096, heads=5
    B, T, H = 2, 4096, 5
    q = torch.randn(B * T, 3 * H * 64, device="cuda", generator=g).half()
    o = torch.empty(B * T, H * 64, dtype=torch.float16, device="cuda")
    ld = 3 * H * 64
    k = q[:, H * 64:]; v = q[:, 2 * H * 64:]
    call = lambda: L.check(L.lib().mi3d_flash_attn_f16(L.ptr(q), C.c_void_p(k.data_ptr()), C.c_void_p(v.data_ptr()), L.ptr(o), C.c_int(B), C.c_int(T), C.c_int(T),
                                                       C.c_int(T), C.c_int(H), C.c_int(ld), C.c_int(ld), C.c_int(ld), C.c_int(H * 64), L.stream()), "attn")
    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"flash attention B={B} T={T} heads={H}: {us:.1f} us/launch, {4.0 * B * H * T * T * 64 / us / 1e6:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
