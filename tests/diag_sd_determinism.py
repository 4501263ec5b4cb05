"""Diagnostic: is the SD engine run-to-run / instance-to-instance / replay deterministic?  (tiny config, prints max rel diffs)"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sd_ref  # noqa: E402


def main():
    sdm = importlib.import_module("make-it-3d_b200.nerf.sd")
    torch.manual_seed(0)
    ucfg, vcfg = sd_ref.tiny_unet_config(), sd_ref.tiny_vae_config()
    unet, vae = sd_ref.UNet2DConditionModel(ucfg).eval(), sd_ref.AutoencoderKLEncoder(vcfg).eval()
    with torch.no_grad():
        for m in list(unet.modules()) + list(vae.modules()):
            if isinstance(m, (torch.nn.GroupNorm, torch.nn.LayerNorm)):
                m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.3, 0.3)
    mk = lambda gr: sdm.StableDiffusion("cuda", unet_cfg=dict(ucfg, latent_hw=32), vae_cfg=dict(vcfg, image_hw=256), unet_state=unet.state_dict(),
                                        vae_state=vae.state_dict(), graph_replay=gr)
    gen = torch.Generator().manual_seed(7)
    lat = (torch.randn(1, 4, 32, 32, generator=gen) * 0.8).cuda()
    noise = torch.randn(1, 4, 32, 32, generator=gen).cuda()
    ctx = torch.randn(2, 77, 128, generator=gen).cuda()
    rgb = torch.rand(1, 3, 64, 64, generator=gen).cuda()
    eps = torch.randn(1, 4, 32, 32, generator=gen).cuda()
    glat = torch.randn(1, 4, 32, 32, generator=gen).cuda()
    tt = torch.tensor([512], dtype=torch.long, device="cuda")
    taps = ["unet.conv_in", "unet.down0", "unet.down1", "unet.down2", "unet.mid", "unet.up0", "unet.up1", "unet.up2"]

    def one(m):
        npred, grad = m.unet_sds(lat, noise, tt, ctx, 10.0)
        torch.cuda.synchronize()
        tp = [m.engine.debug_tensor(n).float().clone() for n in taps]
        r = rgb.clone().requires_grad_()
        z = m.encode_imgs(r, eps)
        z.backward(glat)
        torch.cuda.synchronize()
        return [npred.clone(), grad.clone(), z.detach().clone(), r.grad.clone()] + tp

    def diff(a, b):
        return " ".join(f"{float((x - y).abs().max() / y.abs().max()):.1e}" for x, y in zip(a, b))
    names = "npred grad z dRGB " + " ".join(t.split('.')[1] for t in taps)
    print("columns:", names)
    A = mk(False)
    a1, a2 = one(A), one(A)
    print("plain A run1 vs run2      :", diff(a2, a1))
    B = mk(False)
    b1 = one(B)
    print("plain B (fresh) vs A      :", diff(b1, a1))
    Cg = mk(True)
    for i in range(4):
        c = one(Cg)
        print(f"graph C call {i} vs A       :", diff(c, a1), "replays", Cg.engine.graph_replays())


if __name__ == "__main__":
    main()
