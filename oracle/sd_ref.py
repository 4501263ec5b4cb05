"""oracle/sd_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-PyTorch fp32 restatement of what nerf/sd.py:117-174, 212-220 executes inside the THIRD-PARTY `diffusers` package
(un-pinned git HEAD, README.md:45; NOT in /root/reference, no weights offline):
    UNet2DConditionModel  (SD-2.0-base config: block_out (320,640,1280,1280), 2 layers/block, head_dim 64, cross dim 1024,
                           linear projections, GroupNorm32, SiLU, GEGLU)              -- nerf/sd.py:53,146
    AutoencoderKL.encode + DiagonalGaussianDistribution.sample                       -- nerf/sd.py:41,212-220
    DDIMScheduler.add_noise / alphas_cumprod (scaled-linear betas 0.00085..0.012)     -- nerf/sd.py:55,141,163
plus the reference's own train_step arithmetic (interp, CFG combination, SDS weight, nan_to_num).
Parameter names follow diffusers' state_dict so real checkpoints can be loaded into both this oracle and the product.
PARITY UNPINNED against the real diffusers package (absent here); architecture recalled from its public source.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def sd20_unet_config():
    return dict(in_channels=4, out_channels=4, block_out=(320, 640, 1280, 1280), layers_per_block=2, heads=(5, 10, 20, 20),
                cross_dim=1024, groups=32, ctx_len=77)


def sd_vae_config():
    return dict(in_channels=3, latent_channels=4, block_out=(128, 256, 512, 512), layers_per_block=2, groups=32)


def tiny_unet_config():
    """3-level config used by parity tests (deepest level 8x8 for 32x32 latents)."""
    return dict(in_channels=4, out_channels=4, block_out=(64, 128, 128), layers_per_block=1, heads=(1, 2, 2), cross_dim=128, groups=32,
                ctx_len=77)


def tiny_vae_config():
    return dict(in_channels=3, latent_channels=4, block_out=(64, 64, 128, 128), layers_per_block=1, groups=32)


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_ch, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_ch, cout) if temb_ch else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, dim, heads, cross_dim=None, qkv_bias=False):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=qkv_bias)
        self.to_k = nn.Linear(cross_dim or dim, dim, bias=qkv_bias)
        self.to_v = nn.Linear(cross_dim or dim, dim, bias=qkv_bias)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim)])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        B, T, C = x.shape
        h, d = self.heads, C // self.heads
        q = self.to_q(x).view(B, T, h, d).transpose(1, 2)
        k = self.to_k(ctx).view(B, -1, h, d).transpose(1, 2)
        v = self.to_v(ctx).view(B, -1, h, d).transpose(1, 2)
        p = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5, dim=-1)
        return self.to_out[0]((p @ v).transpose(1, 2).reshape(B, T, C))


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, cross_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, cross_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), ctx) + x
        return self.ff(self.norm3(x)) + x


class Transformer2DModel(nn.Module):
    def __init__(self, dim, heads, cross_dim, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, cross_dim)])
        self.proj_out = nn.Linear(dim, dim)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        res = x
        h = self.norm(x).permute(0, 2, 3, 1).reshape(B, H * W, C)
        h = self.proj_in(h)
        for blk in self.transformer_blocks:
            h = blk(h, ctx)
        h = self.proj_out(h).reshape(B, H, W, C).permute(0, 3, 1, 2)
        return h + res


class Downsample2D(nn.Module):
    def __init__(self, ch, vae=False):
        super().__init__()
        self.vae = vae
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=0 if vae else 1)

    def forward(self, x):
        if self.vae:
            x = F.pad(x, (0, 1, 0, 1))
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.resnets = nn.ModuleList()
        self.attentions = nn.ModuleList()


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        bo, L, g = cfg["block_out"], cfg["layers_per_block"], cfg["groups"]
        temb = bo[0] * 4
        self.conv_in = nn.Conv2d(cfg["in_channels"], bo[0], 3, padding=1)
        self.time_embedding = nn.Module()
        self.time_embedding.linear_1 = nn.Linear(bo[0], temb)
        self.time_embedding.linear_2 = nn.Linear(temb, temb)
        nlev = len(bo)
        self.down_blocks = nn.ModuleList()
        ch = bo[0]
        skip = [ch]
        for i in range(nlev):
            blk = _Block()
            has_attn = i < nlev - 1
            for j in range(L):
                blk.resnets.append(ResnetBlock2D(ch, bo[i], temb, g, 1e-5))
                ch = bo[i]
                if has_attn:
                    blk.attentions.append(Transformer2DModel(ch, cfg["heads"][i], cfg["cross_dim"], g))
                skip.append(ch)
            if i < nlev - 1:
                blk.downsamplers = nn.ModuleList([Downsample2D(ch)])
                skip.append(ch)
            self.down_blocks.append(blk)
        self.mid_block = _Block()
        self.mid_block.resnets.append(ResnetBlock2D(ch, ch, temb, g, 1e-5))
        self.mid_block.attentions.append(Transformer2DModel(ch, cfg["heads"][-1], cfg["cross_dim"], g))
        self.mid_block.resnets.append(ResnetBlock2D(ch, ch, temb, g, 1e-5))
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(bo))
        rev_heads = list(reversed(cfg["heads"]))
        for i in range(nlev):
            blk = _Block()
            has_attn = i > 0
            for j in range(L + 1):
                s = skip.pop()
                blk.resnets.append(ResnetBlock2D(ch + s, rev[i], temb, g, 1e-5))
                ch = rev[i]
                if has_attn:
                    blk.attentions.append(Transformer2DModel(ch, rev_heads[i], cfg["cross_dim"], g))
            if i < nlev - 1:
                blk.upsamplers = nn.ModuleList([Upsample2D(ch)])
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(g, bo[0], eps=1e-5)
        self.conv_out = nn.Conv2d(bo[0], cfg["out_channels"], 3, padding=1)

    def time_proj(self, t):
        half = self.cfg["block_out"][0] // 2
        exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half
        emb = t.float()[:, None] * torch.exp(exponent)[None]
        return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)       # flip_sin_to_cos=True

    def forward(self, x, t, ctx):
        t = torch.as_tensor(t).reshape(-1).expand(x.shape[0])
        temb = self.time_embedding.linear_2(F.silu(self.time_embedding.linear_1(self.time_proj(t))))
        h = self.conv_in(x)
        self.taps = {'unet.conv_in': h}
        skips = [h]
        for bi, blk in enumerate(self.down_blocks):
            for j, res in enumerate(blk.resnets):
                h = res(h, temb)
                if len(blk.attentions):
                    h = blk.attentions[j](h, ctx)
                skips.append(h)
            if hasattr(blk, "downsamplers"):
                h = blk.downsamplers[0](h)
                skips.append(h)
            self.taps[f'unet.down{bi}'] = h
        h = self.mid_block.resnets[0](h, temb)
        h = self.mid_block.attentions[0](h, ctx)
        h = self.mid_block.resnets[1](h, temb)
        self.taps['unet.mid'] = h
        for bi, blk in enumerate(self.up_blocks):
            for j, res in enumerate(blk.resnets):
                h = res(torch.cat([h, skips.pop()], dim=1), temb)
                if len(blk.attentions):
                    h = blk.attentions[j](h, ctx)
            if hasattr(blk, "upsamplers"):
                h = blk.upsamplers[0](h)
            self.taps[f'unet.up{bi}'] = h
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class VaeAttention(nn.Module):
    def __init__(self, ch, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.to_q = nn.Linear(ch, ch)
        self.to_k = nn.Linear(ch, ch)
        self.to_v = nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch)])

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).permute(0, 2, 3, 1).reshape(B, H * W, C)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        p = torch.softmax(q @ k.transpose(-1, -2) * C ** -0.5, dim=-1)
        o = self.to_out[0](p @ v).reshape(B, H, W, C).permute(0, 3, 1, 2)
        return o + x


class Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        bo, L, g = cfg["block_out"], cfg["layers_per_block"], cfg["groups"]
        self.conv_in = nn.Conv2d(cfg["in_channels"], bo[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        ch = bo[0]
        for i in range(len(bo)):
            blk = _Block()
            for j in range(L):
                blk.resnets.append(ResnetBlock2D(ch, bo[i], 0, g, 1e-6))
                ch = bo[i]
            if i < len(bo) - 1:
                blk.downsamplers = nn.ModuleList([Downsample2D(ch, vae=True)])
            self.down_blocks.append(blk)
        self.mid_block = _Block()
        self.mid_block.resnets.append(ResnetBlock2D(ch, ch, 0, g, 1e-6))
        self.mid_block.attentions.append(VaeAttention(ch, g))
        self.mid_block.resnets.append(ResnetBlock2D(ch, ch, 0, g, 1e-6))
        self.conv_norm_out = nn.GroupNorm(g, ch, eps=1e-6)
        self.conv_out = nn.Conv2d(ch, 2 * cfg["latent_channels"], 3, padding=1)

    def forward(self, x):
        h = self.conv_in(x)
        for blk in self.down_blocks:
            for res in blk.resnets:
                h = res(h)
            if hasattr(blk, "downsamplers"):
                h = blk.downsamplers[0](h)
        h = self.mid_block.resnets[0](h)
        h = self.mid_block.attentions[0](h)
        h = self.mid_block.resnets[1](h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class AutoencoderKLEncoder(nn.Module):
    """`vae.encode(x).latent_dist` : encoder + quant_conv -> (mean, logvar)."""

    def __init__(self, cfg):
        super().__init__()
        self.encoder = Encoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg["latent_channels"], 2 * cfg["latent_channels"], 1)

    def forward(self, x):
        mean, logvar = self.quant_conv(self.encoder(x)).chunk(2, dim=1)
        return mean, torch.clamp(logvar, -30.0, 20.0)


def alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
    """DDIMScheduler / SD scheduler_config.json: scaled_linear betas."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def sds_train_step_ref(unet, vae, text_embeddings, pred_rgb, t, eps_posterior, eps_noise, guidance_scale=10.0, alphas=None):
    """nerf/sd.py:117-174 (SDS branch) with the random draws (t, posterior noise, noise) injected.
    Returns dict(latents, latents_noisy, noise_pred (after CFG), grad); performs latents.backward(grad) into pred_rgb's graph."""
    if alphas is None:
        alphas = alphas_cumprod()
    pred_rgb_512 = F.interpolate(pred_rgb, (512, 512), mode='bilinear', align_corners=False) if pred_rgb.shape[-1] != 512 or True else pred_rgb
    return _sds_core(unet, vae, text_embeddings, pred_rgb_512, t, eps_posterior, eps_noise, guidance_scale, alphas)


def sds_train_step_ref_at(unet, vae, text_embeddings, pred_rgb, t, eps_posterior, eps_noise, size, guidance_scale=10.0, alphas=None):
    """Same, but interpolating to (size,size) instead of 512 so tiny test configurations stay cheap."""
    if alphas is None:
        alphas = alphas_cumprod()
    img = F.interpolate(pred_rgb, (size, size), mode='bilinear', align_corners=False)
    return _sds_core(unet, vae, text_embeddings, img, t, eps_posterior, eps_noise, guidance_scale, alphas)


def _sds_core(unet, vae, text_embeddings, img, t, eps_posterior, eps_noise, guidance_scale, alphas):
    mean, logvar = vae(2 * img - 1)                                      # sd.py:215-217
    latents = (mean + torch.exp(0.5 * logvar) * eps_posterior) * 0.18215  # sd.py:218
    with torch.no_grad():
        a = alphas[t]
        latents_noisy = a.sqrt() * latents + (1 - a).sqrt() * eps_noise   # DDIMScheduler.add_noise, sd.py:141
        x = torch.cat([latents_noisy] * 2)
        noise_pred = unet(x, torch.tensor([t]), text_embeddings)          # sd.py:146
        uncond, text = noise_pred.chunk(2)
        noise_pred = text + guidance_scale * (text - uncond)              # sd.py:150-151 (NB: text + gs*(text-uncond))
        w = 1 - a
        grad = torch.nan_to_num(w * (noise_pred - eps_noise))             # sd.py:165-170
    if latents.requires_grad:
        latents.backward(gradient=grad, retain_graph=True)                # sd.py:171
    return dict(latents=latents.detach(), latents_noisy=latents_noisy, noise_pred=noise_pred, grad=grad)


# ------------------------------------------------------------------------------------------------------------
# The "denoise" side branch of train_step (nerf/sd.py:153-159): one DDIM step t -> t-1 and the VAE decoder (nerf/sd.py:201-210).
# ------------------------------------------------------------------------------------------------------------
class Decoder(nn.Module):
    """diffusers `Decoder` of AutoencoderKL: conv_in, mid (res, attention, res), UpDecoderBlock2D x len(block_out) with
    layers_per_block + 1 resnets each and nearest-2x + conv upsamplers on all but the last, GroupNorm + SiLU + conv_out."""

    def __init__(self, cfg):
        super().__init__()
        bo, L, g = cfg["block_out"], cfg["layers_per_block"], cfg["groups"]
        rev = list(reversed(bo))
        ch = rev[0]
        self.conv_in = nn.Conv2d(cfg["latent_channels"], ch, 3, padding=1)
        self.mid_block = _Block()
        self.mid_block.resnets.append(ResnetBlock2D(ch, ch, 0, g, 1e-6))
        self.mid_block.attentions.append(VaeAttention(ch, g))
        self.mid_block.resnets.append(ResnetBlock2D(ch, ch, 0, g, 1e-6))
        self.up_blocks = nn.ModuleList()
        for i, co in enumerate(rev):
            blk = _Block()
            for j in range(L + 1):
                blk.resnets.append(ResnetBlock2D(ch, co, 0, g, 1e-6))
                ch = co
            if i < len(rev) - 1:
                blk.upsamplers = nn.ModuleList([Upsample2D(ch)])
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(g, ch, eps=1e-6)
        self.conv_out = nn.Conv2d(ch, cfg["in_channels"], 3, padding=1)

    def forward(self, z):
        h = self.conv_in(z)
        h = self.mid_block.resnets[0](h)
        h = self.mid_block.attentions[0](h)
        h = self.mid_block.resnets[1](h)
        for blk in self.up_blocks:
            for res in blk.resnets:
                h = res(h)
            if hasattr(blk, "upsamplers"):
                h = blk.upsamplers[0](h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class AutoencoderKLDecoder(nn.Module):
    """`vae.decode(z).sample`: post_quant_conv (1x1) + decoder; parameter names as in AutoencoderKL.state_dict()."""

    def __init__(self, cfg):
        super().__init__()
        self.post_quant_conv = nn.Conv2d(cfg["latent_channels"], cfg["latent_channels"], 1)
        self.decoder = Decoder(cfg)

    def forward(self, z):
        return self.decoder(self.post_quant_conv(z))


def ddim_step_ref(noise_pred, t, latents_noisy, alphas=None):
    """DDIMScheduler.step(model_output, t, sample) with eta = 0 after set_timesteps(num_train_timesteps) (nerf/sd.py:154-155):
    prev = t - 1; alpha_prev = alphas_cumprod[prev] if prev >= 0 else alphas_cumprod[0] (set_alpha_to_one=False in SD's scheduler
    config); epsilon prediction, no clipping: x0 = (x_t - sqrt(1 - a_t) eps) / sqrt(a_t); x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev) eps."""
    if alphas is None:
        alphas = alphas_cumprod()
    a_t = alphas[t]
    a_prev = alphas[t - 1] if t - 1 >= 0 else alphas[0]
    x0 = (latents_noisy - (1 - a_t).sqrt() * noise_pred) / a_t.sqrt()
    return a_prev.sqrt() * x0 + (1 - a_prev).sqrt() * noise_pred


def decode_latents_ref(vae_dec, latents):
    """nerf/sd.py:201-210"""
    with torch.no_grad():
        imgs = vae_dec(latents / 0.18215)
    return (imgs / 2 + 0.5).clamp(0, 1)
