"""Oracle: Stable-Diffusion-v1 UNet + KL-VAE as used by ODISE (CPU, plain PyTorch fp32/fp64).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED for the module internals: the arithmetic lives in `ldm` from stable-diffusion-sdkit==2.1.3
(reference setup.py:87; import sites odise/modeling/meta_arch/ldm.py:17-20), which is NOT vendored in
/root/reference and is not installed here.  This file restates the published algorithm of
ldm/modules/diffusionmodules/openaimodel.py (UNetModel, ResBlock, Upsample, Downsample, timestep_embedding),
ldm/modules/attention.py (SpatialTransformer, BasicTransformerBlock, CrossAttention, GEGLU, FeedForward),
ldm/modules/diffusionmodules/model.py (Encoder, Decoder, ResnetBlock, AttnBlock) with the v1-inference.yaml
hyper-parameters, using ldm's parameter names so real sd-v1 checkpoints stay loadable.
What IS pinned: the attribute surface the reference touches — tests/test_oracle_cpu.py runs the
reference's own LdmExtractor.unet_forward / encoder_forward / decoder_forward (ldm.py:424-533) verbatim on top of
these modules when /root/reference is present, and checks the tap channel/stride table of ldm.py:284-346.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def timestep_embedding(timesteps, dim, max_period=10000, repeat_only=False):
    """openaimodel/util.timestep_embedding: cat(cos, sin) (cos first)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


class GroupNorm32(nn.GroupNorm):
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


class TimestepEmbedSequential(nn.Sequential):
    def forward(self, x, emb, context=None):
        for layer in self:
            if isinstance(layer, ResBlock):
                x = layer(x, emb)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context)
            else:
                x = layer(x)
        return x


class Upsample(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.channels = channels
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2, mode="nearest"))


class Downsample(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.channels = channels
        self.op = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.op(x)


class ResBlock(nn.Module):
    def __init__(self, channels, emb_channels, out_channels=None):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.in_layers = nn.Sequential(GroupNorm32(32, channels), nn.SiLU(),
                                       nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(GroupNorm32(32, self.out_channels), nn.SiLU(), nn.Dropout(0.0),
                                        nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 1)

    def forward(self, x, emb):
        h = self.in_layers(x)
        emb_out = self.emb_layers(emb).type(h.dtype)
        while len(emb_out.shape) < len(h.shape):
            emb_out = emb_out[..., None]
        h = h + emb_out
        h = self.out_layers(h)
        return self.skip_connection(x) + h


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = dim_head * heads
        context_dim = context_dim or query_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(0.0))

    def forward(self, x, context=None):
        h = self.heads
        q = self.to_q(x)
        context = x if context is None else context
        k, v = self.to_k(context), self.to_v(context)
        b, n, _ = q.shape
        q, k, v = (t.reshape(b, t.shape[1], h, -1).permute(0, 2, 1, 3) for t in (q, k, v))
        sim = torch.einsum("bhid,bhjd->bhij", q, k) * self.scale
        attn = sim.softmax(dim=-1)
        out = torch.einsum("bhij,bhjd->bhid", attn, v)
        out = out.permute(0, 2, 1, 3).reshape(b, n, -1)
        return self.to_out(out)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.Sequential(GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim))

    def forward(self, x):
        return self.net(x)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, n_heads, d_head, context_dim):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, n_heads, d_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim, n_heads, d_head)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)

    def forward(self, x, context=None):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context=context) + x
        x = self.ff(self.norm3(x)) + x
        return x


class SpatialTransformer(nn.Module):
    def __init__(self, in_channels, n_heads, d_head, context_dim):
        super().__init__()
        inner = n_heads * d_head
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, n_heads, d_head, context_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, context=None):
        b, c, h, w = x.shape
        x_in = x
        x = self.proj_in(self.norm(x))
        x = x.flatten(2).transpose(1, 2)
        for blk in self.transformer_blocks:
            x = blk(x, context=context)
        x = x.transpose(1, 2).reshape(b, c, h, w)
        return self.proj_out(x) + x_in


class UNetModel(nn.Module):
    """SD-v1: in 4, model_channels 320, mult (1,2,4,4), 2 res blocks, attention at ds 1,2,4, 8 heads, ctx 768."""

    def __init__(self, in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2,
                 attention_resolutions=(4, 2, 1), channel_mult=(1, 2, 4, 4), num_heads=8, context_dim=768):
        super().__init__()
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        ted = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, ted, mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(SpatialTransformer(ch, num_heads, ch // num_heads, context_dim))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(
            ResBlock(ch, ted), SpatialTransformer(ch, num_heads, ch // num_heads, context_dim), ResBlock(ch, ted))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [ResBlock(ch + ich, ted, model_channels * mult)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(SpatialTransformer(ch, num_heads, ch // num_heads, context_dim))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(GroupNorm32(32, ch), nn.SiLU(), nn.Conv2d(model_channels, out_channels, 3, padding=1))

    def forward(self, x, timesteps, context):
        hs = []
        emb = self.time_embed(timestep_embedding(timesteps, self.model_channels))
        h = x
        for m in self.input_blocks:
            h = m(h, emb, context)
            hs.append(h)
        h = self.middle_block(h, emb, context)
        for m in self.output_blocks:
            h = m(torch.cat([h, hs.pop()], dim=1), emb, context)
        return self.out(h)


# ----------------------------------------------------------------------------------------------- VAE (model.py)
def Normalize(c):
    return nn.GroupNorm(32, c, eps=1e-6, affine=True)


def nonlinearity(x):
    return x * torch.sigmoid(x)


class ResnetBlock(nn.Module):
    def __init__(self, in_channels, out_channels=None):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels or in_channels
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, self.out_channels, 3, padding=1)
        self.norm2 = Normalize(self.out_channels)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1)
        if self.in_channels != self.out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, self.out_channels, 1)

    def forward(self, x, temb=None):
        h = self.conv1(nonlinearity(self.norm1(x)))
        h = self.conv2(self.dropout(nonlinearity(self.norm2(h))))
        if self.in_channels != self.out_channels:
            x = self.nin_shortcut(x)
        return x + h


class AttnBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.in_channels = c
        self.norm = Normalize(c)
        self.q, self.k, self.v = nn.Conv2d(c, c, 1), nn.Conv2d(c, c, 1), nn.Conv2d(c, c, 1)
        self.proj_out = nn.Conv2d(c, c, 1)

    def forward(self, x):
        h_ = self.norm(x)
        q, k, v = self.q(h_), self.k(h_), self.v(h_)
        b, c, h, w = q.shape
        q = q.reshape(b, c, h * w).permute(0, 2, 1)
        k = k.reshape(b, c, h * w)
        w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
        w_ = F.softmax(w_, dim=2)
        v = v.reshape(b, c, h * w)
        h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
        return x + self.proj_out(h_)


class VaeDownsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class VaeUpsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Level(nn.Module):
    pass


class Encoder(nn.Module):
    def __init__(self, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, z_channels=4, double_z=True):
        super().__init__()
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.conv_in = nn.Conv2d(in_channels, ch, 3, padding=1)
        in_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        for i in range(self.num_resolutions):
            lvl = _Level()
            lvl.block, lvl.attn = nn.ModuleList(), nn.ModuleList()
            bi, bo = ch * in_mult[i], ch * ch_mult[i]
            for _ in range(num_res_blocks):
                lvl.block.append(ResnetBlock(bi, bo))
                bi = bo
            if i != self.num_resolutions - 1:
                lvl.downsample = VaeDownsample(bi)
            self.down.append(lvl)
        self.mid = _Level()
        self.mid.block_1, self.mid.attn_1, self.mid.block_2 = ResnetBlock(bi), AttnBlock(bi), ResnetBlock(bi)
        self.norm_out = Normalize(bi)
        self.conv_out = nn.Conv2d(bi, 2 * z_channels if double_z else z_channels, 3, padding=1)

    def forward(self, x):
        h = self.conv_in(x)
        for i in range(self.num_resolutions):
            for j in range(self.num_res_blocks):
                h = self.down[i].block[j](h, None)
            if i != self.num_resolutions - 1:
                h = self.down[i].downsample(h)
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h, None)), None)
        return self.conv_out(nonlinearity(self.norm_out(h)))


class Decoder(nn.Module):
    def __init__(self, ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4):
        super().__init__()
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.give_pre_end, self.tanh_out, self.last_z_shape = False, False, None
        bi = ch * ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, bi, 3, padding=1)
        self.mid = _Level()
        self.mid.block_1, self.mid.attn_1, self.mid.block_2 = ResnetBlock(bi), AttnBlock(bi), ResnetBlock(bi)
        self.up = nn.ModuleList()
        for i in reversed(range(self.num_resolutions)):
            lvl = _Level()
            lvl.block, lvl.attn = nn.ModuleList(), nn.ModuleList()
            bo = ch * ch_mult[i]
            for _ in range(num_res_blocks + 1):
                lvl.block.append(ResnetBlock(bi, bo))
                bi = bo
            if i != 0:
                lvl.upsample = VaeUpsample(bi)
            self.up.insert(0, lvl)
        self.norm_out = Normalize(bi)
        self.conv_out = nn.Conv2d(bi, out_ch, 3, padding=1)

    def forward(self, z):
        h = self.conv_in(z)
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h, None)), None)
        for i in reversed(range(self.num_resolutions)):
            for j in range(self.num_res_blocks + 1):
                h = self.up[i].block[j](h, None)
            if i != 0:
                h = self.up[i].upsample(h)
        return self.conv_out(nonlinearity(self.norm_out(h)))


class AutoencoderKL(nn.Module):
    def __init__(self, embed_dim=4, z_channels=4):
        super().__init__()
        self.embed_dim = embed_dim
        self.encoder, self.decoder = Encoder(z_channels=z_channels), Decoder(z_channels=z_channels)
        self.quant_conv = nn.Conv2d(2 * z_channels, 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, z_channels, 1)


class DiagonalGaussianDistribution:
    def __init__(self, parameters):
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)


# ----------------------------------------------------------------------------------------------- ODISE driver
# constants of GaussianDiffusion.q_sample at t = 0 with the "ldm_linear" schedule
# (odise/modeling/diffusion/gaussian_diffusion.py:125-135, :275-292): beta_0 = (sqrt(0.00085))^2
def _t0_coeffs():
    import numpy as np
    betas = np.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=np.float64) ** 2   # "ldm_linear", steps=1000
    ac = np.cumprod(1.0 - betas, axis=0)
    # _extract_into_tensor(...).float(): the fp64 table entry is rounded to fp32 before the multiply
    return float(np.float32(np.sqrt(ac)[0])), float(np.float32(np.sqrt(1.0 - ac)[0]))


SQRT_ALPHA_BAR_0, SQRT_ONE_MINUS_ALPHA_BAR_0 = _t0_coeffs()
SCALE_FACTOR = 0.18215
UNET_TAP_BLOCKS = (2, 5, 8, 11)   # configs/common/models/odise_with_label.py:20
ENC_TAP_BLOCKS = (5, 7)
DEC_TAP_BLOCKS = (2, 5)


def shared_noise(latent_hw=(64, 64)):
    """LdmExtractor.__init__ (ldm.py:271-277): randn(1, 4, 64, 64) from a CPU generator seeded with 42."""
    rng = torch.Generator().manual_seed(42)
    n = torch.randn(1, 4, 64, 64, generator=rng)
    if tuple(latent_hw) != (64, 64):   # ldm.py:585-591
        n = F.interpolate(n, size=latent_hw, mode="bicubic", align_corners=False)
    return n


def unet_features(unet, x, context, cond_emb=None, tap_blocks=UNET_TAP_BLOCKS, stop_early=True):
    """LdmExtractor.unet_forward (ldm.py:469-491) at t = 0: returns the inputs of the tapped output blocks.
    stop_early=False also executes output block 11 and unet.out like the reference does (results discarded,
    ldm.py:600) — used when timing the reference's CPU path."""
    t = torch.zeros(x.shape[0], dtype=torch.long, device=x.device)
    emb = unet.time_embed(timestep_embedding(t, unet.model_channels))
    if cond_emb is not None:
        emb = emb + cond_emb
    hs, feats = [], []
    h = x
    for m in unet.input_blocks:
        h = m(h, emb, context)
        hs.append(h)
    h = unet.middle_block(h, emb, context)
    for i, m in enumerate(unet.output_blocks):
        h = torch.cat([h, hs.pop()], dim=1)
        if i in tap_blocks:
            feats.append(h.contiguous())
        if stop_early and i == max(tap_blocks):
            break   # remaining work is dead code for ODISE
        h = m(h, emb, context)
    if not stop_early:
        unet.out(h)
    return feats


def encoder_features(vae, x, tap_blocks=ENC_TAP_BLOCKS):
    """LdmExtractor.encoder_forward + encode_to_latent (ldm.py:424-467): latent = 0.18215 * posterior mean."""
    enc = vae.encoder
    feats = []
    h = enc.conv_in(x)
    idx = 0
    for i in range(enc.num_resolutions):
        for j in range(enc.num_res_blocks):
            if idx in tap_blocks:
                feats.append(h.contiguous())
            h = enc.down[i].block[j](h, None)
            idx += 1
        if i != enc.num_resolutions - 1:
            h = enc.down[i].downsample(h)
    h = enc.mid.block_2(enc.mid.attn_1(enc.mid.block_1(h, None)), None)
    h = enc.conv_out(nonlinearity(enc.norm_out(h)))
    moments = vae.quant_conv(h)
    return SCALE_FACTOR * DiagonalGaussianDistribution(moments).mean, feats


def decoder_features(vae, latent, tap_blocks=DEC_TAP_BLOCKS, truncate=True):
    """LdmExtractor.decode_to_image / decoder_forward (ldm.py:493-541) truncated after the last tap
    (truncate=False runs on to the full RGB image like the reference, for CPU timing)."""
    dec = vae.decoder
    z = vae.post_quant_conv(1.0 / SCALE_FACTOR * latent)   # ldm.py:536
    h = dec.conv_in(z)
    h = dec.mid.block_2(dec.mid.attn_1(dec.mid.block_1(h, None)), None)
    feats = []
    idx = 0
    for i in reversed(range(dec.num_resolutions)):
        for j in range(dec.num_res_blocks + 1):
            if idx in tap_blocks:
                feats.append(h.contiguous())
                if truncate and idx == max(tap_blocks):
                    return feats
            h = dec.up[i].block[j](h, None)
            idx += 1
        if i != 0:
            h = dec.up[i].upsample(h)
    if not truncate:
        dec.conv_out(nonlinearity(dec.norm_out(h)))
    return feats


def q_sample_t0(latent, noise):
    """GaussianDiffusion.q_sample at t=0 (gaussian_diffusion.py:275-292)."""
    return SQRT_ALPHA_BAR_0 * latent + SQRT_ONE_MINUS_ALPHA_BAR_0 * noise.expand_as(latent)
