"""GPU exploration: which MN-major descriptor conventions produce correct results (prints errors, never asserts)."""
import ctypes as C, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module("make-it-3d_b200._lib")
which = sys.argv[1]
g = torch.Generator(device="cuda").manual_seed(0)
if which == "f16bt":
    for (M, N, K, bn) in [(128, 64, 64, 64), (128, 64, 256, 64), (256, 128, 128, 128), (128, 256, 64, 256)]:
        a = torch.randn(M, K, device="cuda", generator=g).half(); bt = (torch.randn(K, N, device="cuda", generator=g) / K ** 0.5).half()
        out = torch.zeros(M, N, device="cuda")
        r = L.lib().mi3d_gemm_f16_bt(L.ptr(a), L.ptr(bt), L.ptr(out), C.c_int(M), C.c_int(N), C.c_int(K), C.c_int(bn), L.stream())
        torch.cuda.synchronize()
        ref = a.float() @ bt.float()
        print("f16 B MN-major", (M, N, K, bn), "rc", r, "max err", (out - ref).abs().max().item(), "ref max", ref.abs().max().item(), flush=True)
else:
    mode, N, K = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    m = mode - 2 if mode >= 3 else mode
    if m == 0:
        a = torch.randn(128, K, device="cuda", generator=g); b = torch.randn(N, K, device="cuda", generator=g); ref = a.double() @ b.double().t()
    elif m == 1:
        a = torch.randn(K, 128, device="cuda", generator=g); b = torch.randn(K, N, device="cuda", generator=g); ref = a.double().t() @ b.double()
    else:
        a = torch.randn(128, K, device="cuda", generator=g); b = torch.randn(K, N, device="cuda", generator=g); ref = a.double() @ b.double()
    d = torch.zeros(128, N, device="cuda")
    r = L.lib().mi3d_tf32_tile_test(L.ptr(a), L.ptr(b), L.ptr(d), C.c_int(N), C.c_int(K), C.c_int(mode), L.stream())
    torch.cuda.synchronize()
    print("tf32 tile mode", mode, (N, K), "rc", r, "max err", (d.double() - ref).abs().max().item(), "ref max", ref.abs().max().item(), flush=True)
