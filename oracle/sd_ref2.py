"""oracle/sd_ref2.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A SECOND, independently written restatement of the diffusers arithmetic behind nerf/sd.py:146 (`UNet2DConditionModel`) and
nerf/sd.py:217 (`AutoencoderKL.encode`): purely FUNCTIONAL, driven by a diffusers-named state_dict (no module classes, no config:
the topology is read off the parameter names and shapes).  It exists because the real `diffusers` package is absent offline
(un-pinned git HEAD, README.md:45): parity of oracle/sd_ref.py is pinned (a) block by block against torch.nn / torch.nn.functional
building blocks, which both restatements call but compose differently -- sd_ref.py through nn.Module trees, this file through
F.conv2d / F.group_norm / F.layer_norm / F.linear / F.scaled_dot_product_attention on raw tensors, NCHW <-> token reshapes written
the other way round --, (b) against the published parameter counts, and (c) if a real diffusers install ever is available,
against fixtures dumped by tools/dump_diffusers_fixtures.py (tests/test_oracle_golden.py::test_sd_oracle_vs_diffusers_fixtures).
Only tests/ may import this.
"""
import math
import re

import torch
import torch.nn.functional as F


def _count(sd, pattern):
    idx = set()
    for k in sd:
        m = re.match(pattern, k)
        if m:
            idx.add(int(m.group(1)))
    return len(idx)


def _gn_silu(x, sd, p, groups, eps):
    return F.silu(F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps))


def _conv(x, sd, p, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def _resnet(x, sd, p, groups, eps, temb=None):
    h = _conv(_gn_silu(x, sd, p + ".norm1", groups, eps), sd, p + ".conv1")
    if temb is not None:
        h = h + F.linear(F.silu(temb), sd[p + ".time_emb_proj.weight"], sd[p + ".time_emb_proj.bias"])[:, :, None, None]
    h = _conv(_gn_silu(h, sd, p + ".norm2", groups, eps), sd, p + ".conv2")
    if p + ".conv_shortcut.weight" in sd:
        x = _conv(x, sd, p + ".conv_shortcut", padding=0)
    return x + h


def _mha(q_in, kv_in, sd, p, head_dim=64):
    q = F.linear(q_in, sd[p + ".to_q.weight"], sd.get(p + ".to_q.bias"))
    k = F.linear(kv_in, sd[p + ".to_k.weight"], sd.get(p + ".to_k.bias"))
    v = F.linear(kv_in, sd[p + ".to_v.weight"], sd.get(p + ".to_v.bias"))
    B, T, Cc = q.shape
    h = Cc // head_dim
    split = lambda t: t.reshape(B, t.shape[1], h, head_dim).permute(0, 2, 1, 3)
    o = F.scaled_dot_product_attention(split(q), split(k), split(v))            # softmax(q k^T / sqrt(64)) v
    o = o.permute(0, 2, 1, 3).reshape(B, T, Cc)
    return F.linear(o, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def _transformer(x, ctx, sd, p, groups):
    B, Cc, H, W = x.shape
    tok = F.group_norm(x, groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6).flatten(2).transpose(1, 2)     # [B, HW, C]
    tok = F.linear(tok, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    for b in range(_count(sd, re.escape(p) + r"\.transformer_blocks\.(\d+)\.")):
        q = f"{p}.transformer_blocks.{b}"
        ln = lambda t, n: F.layer_norm(t, (Cc,), sd[f"{q}.{n}.weight"], sd[f"{q}.{n}.bias"], 1e-5)
        n1 = ln(tok, "norm1")
        tok = tok + _mha(n1, n1, sd, q + ".attn1")
        tok = tok + _mha(ln(tok, "norm2"), ctx, sd, q + ".attn2")
        a, gate = F.linear(ln(tok, "norm3"), sd[q + ".ff.net.0.proj.weight"], sd[q + ".ff.net.0.proj.bias"]).chunk(2, dim=-1)
        tok = tok + F.linear(a * F.gelu(gate), sd[q + ".ff.net.2.weight"], sd[q + ".ff.net.2.bias"])
    tok = F.linear(tok, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return x + tok.transpose(1, 2).reshape(B, Cc, H, W)


def unet_forward(sd, x, t, ctx, groups=32):
    """UNet2DConditionModel.forward(x, t, encoder_hidden_states=ctx).sample for SD-2.x style configs (cross-attention in every
    level but the deepest, linear projections, GEGLU, head_dim 64, flip_sin_to_cos, freq_shift 0)."""
    c0 = sd["conv_in.weight"].shape[0]
    half = c0 // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    ang = torch.as_tensor(t, dtype=torch.float32).reshape(-1, 1).expand(x.shape[0], 1) * freqs[None]
    temb = torch.cat([ang.cos(), ang.sin()], dim=1)
    temb = F.linear(F.silu(F.linear(temb, sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])),
                    sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])
    h = _conv(x, sd, "conv_in")
    stack = [h]
    n_down = _count(sd, r"down_blocks\.(\d+)\.")
    for i in range(n_down):
        for j in range(_count(sd, rf"down_blocks\.{i}\.resnets\.(\d+)\.")):
            h = _resnet(h, sd, f"down_blocks.{i}.resnets.{j}", groups, 1e-5, temb)
            if f"down_blocks.{i}.attentions.{j}.norm.weight" in sd:
                h = _transformer(h, ctx, sd, f"down_blocks.{i}.attentions.{j}", groups)
            stack.append(h)
        if f"down_blocks.{i}.downsamplers.0.conv.weight" in sd:
            h = _conv(h, sd, f"down_blocks.{i}.downsamplers.0.conv", stride=2)
            stack.append(h)
    h = _resnet(h, sd, "mid_block.resnets.0", groups, 1e-5, temb)
    h = _transformer(h, ctx, sd, "mid_block.attentions.0", groups)
    h = _resnet(h, sd, "mid_block.resnets.1", groups, 1e-5, temb)
    for i in range(_count(sd, r"up_blocks\.(\d+)\.")):
        for j in range(_count(sd, rf"up_blocks\.{i}\.resnets\.(\d+)\.")):
            h = _resnet(torch.cat([h, stack.pop()], dim=1), sd, f"up_blocks.{i}.resnets.{j}", groups, 1e-5, temb)
            if f"up_blocks.{i}.attentions.{j}.norm.weight" in sd:
                h = _transformer(h, ctx, sd, f"up_blocks.{i}.attentions.{j}", groups)
        if f"up_blocks.{i}.upsamplers.0.conv.weight" in sd:
            h = _conv(h.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3), sd, f"up_blocks.{i}.upsamplers.0.conv")   # nearest x2
    assert not stack
    return _conv(_gn_silu(h, sd, "conv_norm_out", groups, 1e-5), sd, "conv_out")


def vae_encode_moments(sd, x, groups=32):
    """AutoencoderKL.encode(x).latent_dist -> (mean, clamped logvar); names as in AutoencoderKL.state_dict()."""
    h = _conv(x, sd, "encoder.conv_in")
    for i in range(_count(sd, r"encoder\.down_blocks\.(\d+)\.")):
        for j in range(_count(sd, rf"encoder\.down_blocks\.{i}\.resnets\.(\d+)\.")):
            h = _resnet(h, sd, f"encoder.down_blocks.{i}.resnets.{j}", groups, 1e-6)
        if f"encoder.down_blocks.{i}.downsamplers.0.conv.weight" in sd:
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"],
                         sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"], stride=2)          # asymmetric pad, no conv padding
    h = _resnet(h, sd, "encoder.mid_block.resnets.0", groups, 1e-6)
    p = "encoder.mid_block.attentions.0"
    B, Cc, H, W = h.shape
    tok = F.group_norm(h, groups, sd[p + ".group_norm.weight"], sd[p + ".group_norm.bias"], 1e-6).flatten(2).transpose(1, 2)
    att = _mha(tok, tok, sd, p, head_dim=Cc)                                                        # single head over all channels
    h = h + att.transpose(1, 2).reshape(B, Cc, H, W)
    h = _resnet(h, sd, "encoder.mid_block.resnets.1", groups, 1e-6)
    h = _conv(_gn_silu(h, sd, "encoder.conv_norm_out", groups, 1e-6), sd, "encoder.conv_out")
    mean, logvar = _conv(h, sd, "quant_conv", padding=0).chunk(2, dim=1)
    return mean, logvar.clamp(-30.0, 20.0)


def _vae_mid(h, sd, p, groups):
    """mid block of the VAE encoder / decoder: resnet, single-head attention over all channels, resnet"""
    h = _resnet(h, sd, p + ".resnets.0", groups, 1e-6)
    a = p + ".attentions.0"
    B, Cc, H, W = h.shape
    tok = F.group_norm(h, groups, sd[a + ".group_norm.weight"], sd[a + ".group_norm.bias"], 1e-6).flatten(2).transpose(1, 2)
    h = h + _mha(tok, tok, sd, a, head_dim=Cc).transpose(1, 2).reshape(B, Cc, H, W)
    return _resnet(h, sd, p + ".resnets.1", groups, 1e-6)


def vae_decode(sd, z, groups=32):
    """AutoencoderKL.decode(z).sample (nerf/sd.py:205): post_quant_conv (1x1), decoder.conv_in, mid block, up blocks (every resnet of a
    block, then nearest-2x + conv where the block has an upsampler), GroupNorm + SiLU + conv_out; names as in AutoencoderKL.state_dict()."""
    h = _conv(_conv(z, sd, "post_quant_conv", padding=0), sd, "decoder.conv_in")
    h = _vae_mid(h, sd, "decoder.mid_block", groups)
    for i in range(_count(sd, r"decoder\.up_blocks\.(\d+)\.")):
        for j in range(_count(sd, rf"decoder\.up_blocks\.{i}\.resnets\.(\d+)\.")):
            h = _resnet(h, sd, f"decoder.up_blocks.{i}.resnets.{j}", groups, 1e-6)
        if f"decoder.up_blocks.{i}.upsamplers.0.conv.weight" in sd:
            h = _conv(F.interpolate(h, scale_factor=2.0, mode="nearest"), sd, f"decoder.up_blocks.{i}.upsamplers.0.conv")
    return _conv(_gn_silu(h, sd, "decoder.conv_norm_out", groups, 1e-6), sd, "decoder.conv_out")


def ddim_prev_sample(eps, t, x_t, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
    """DDIMScheduler.step (eta = 0, epsilon prediction, no clipping, set_alpha_to_one = False) for prev = t - 1, written from the
    closed form  x_{t-1} = sqrt(a_prev / a_t) x_t + (sqrt(1 - a_prev) - sqrt(a_prev (1 - a_t) / a_t)) eps  in float64."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
    ac = torch.cumprod(1.0 - betas, 0)
    a_t, a_p = ac[t], ac[max(t - 1, 0)]
    c_x = math.sqrt(float(a_p / a_t))
    c_e = math.sqrt(float(1 - a_p)) - math.sqrt(float(a_p * (1 - a_t) / a_t))
    return (c_x * x_t.double() + c_e * eps.double()).to(x_t.dtype)
