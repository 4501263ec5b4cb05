"""Run on a machine that HAS the `diffusers` package (this image does not: no network, package absent -- README.md:45 installs it
from git HEAD).  Builds diffusers' own UNet2DConditionModel / AutoencoderKL / DDIMScheduler at the reduced test configuration,
seeds them, and writes tests/golden/diffusers_tiny.npz: parameters (diffusers state_dict names), inputs and outputs.
tests/test_oracle_golden.py::test_sd_oracle_vs_diffusers_fixtures then pins oracle/sd_ref.py against the real package.

    python tools/dump_diffusers_fixtures.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    try:
        from diffusers import AutoencoderKL, DDIMScheduler, UNet2DConditionModel
    except ImportError:
        print("diffusers is not installed here; nothing written", file=sys.stderr)
        return 1
    torch.manual_seed(0)
    # mirrors oracle/sd_ref.py::tiny_unet_config / tiny_vae_config in diffusers' config vocabulary (SD-2.x style blocks)
    unet = UNet2DConditionModel(sample_size=16, in_channels=4, out_channels=4, block_out_channels=(64, 128, 128), layers_per_block=1,
                                down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
                                up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"), cross_attention_dim=128,
                                attention_head_dim=(1, 2, 2), use_linear_projection=True, norm_num_groups=32).eval()
    vae = AutoencoderKL(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(64, 64, 128, 128), layers_per_block=1,
                        down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4, norm_num_groups=32).eval()
    sched = DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000,
                          clip_sample=False, set_alpha_to_one=False, steps_offset=1)
    g = torch.Generator().manual_seed(1)
    x, ctx = torch.randn(2, 4, 16, 16, generator=g), torch.randn(2, 77, 128, generator=g)
    img = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    t = 437
    with torch.no_grad():
        out = unet(x, t, encoder_hidden_states=ctx).sample
        dist = vae.encode(img).latent_dist
    z = {"in.x": x.numpy(), "in.ctx": ctx.numpy(), "in.img": img.numpy(), "in.t": np.int64(t), "out.unet": out.numpy(),
         "out.mean": dist.mean.numpy(), "out.logvar": dist.logvar.numpy(), "out.alphas_cumprod": sched.alphas_cumprod.numpy()}
    for k, v in unet.state_dict().items():
        z["unet." + k] = v.numpy()
    for k, v in vae.state_dict().items():
        if k.startswith("encoder.") or k.startswith("quant_conv."):
            z["vae." + k] = v.numpy()
    path = os.path.join(ROOT, "tests", "golden", "diffusers_tiny.npz")
    np.savez_compressed(path, **z)
    print("wrote", path)
    return 0


if __name__ == "__main__":
    sys.exit(main())
