"""tests/golden/make_golden_adan.py -- golden vectors of the parameter update from the REFERENCE's own optimizer
(/root/reference/optimizer.py Adan, built exactly like main.py:132, driven exactly like nerf/utils.py:983-986:
clip_grad_norm(max_norm=10) then optimizer.step()) on a small two-group problem shaped like NeRFNetwork.get_params
(group 0 = "encoder" at 10x lr, group 1 = three "MLP" tensors).  Gradient scales are chosen so that step 1 triggers neither
clip, step 2 only Adan's 5.0 clip, steps 3-4 both.  Run in the build container:  python tests/golden/make_golden_adan.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from optimizer import Adan      # noqa: E402  (REFERENCE class)


def main():
    g = torch.Generator().manual_seed(0)
    shapes = [(8192,), (64, 32), (64,), (4, 64)]
    params = [torch.nn.Parameter(torch.randn(s, generator=g) * 0.1) for s in shapes]
    lr = 5 * 1e-3
    opt = Adan([{"params": params[:1], "lr": lr * 10}, {"params": params[1:], "lr": lr}], eps=1e-8, weight_decay=2e-5, max_grad_norm=5.0, foreach=False)
    out = {"lr": np.float64(lr)}
    for i, p in enumerate(params):
        out[f"p0_{i}"] = p.detach().numpy().copy()
    scales = [1e-2, 0.08, 0.5, 3.0]
    for step, sc in enumerate(scales):
        for i, p in enumerate(params):
            p.grad = torch.randn(p.shape, generator=g) * sc
            out[f"g{step}_{i}"] = p.grad.numpy().copy()
        norm = torch.nn.utils.clip_grad_norm_(params, max_norm=10)
        out[f"norm{step}"] = np.float32(norm)
        opt.step()
        for i, p in enumerate(params):
            out[f"p{step + 1}_{i}"] = p.detach().numpy().copy()
            st = opt.state[p]
            for k in ("exp_avg", "exp_avg_sq", "exp_avg_diff", "neg_pre_grad"):
                out[f"{k}{step + 1}_{i}"] = st[k].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "adan.npz"), **out)
    print("wrote adan.npz; norms", [float(out[f"norm{s}"]) for s in range(len(scales))])


if __name__ == "__main__":
    main()
