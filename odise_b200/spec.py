"""Parameter inventory (names + shapes, in the reference's state-dict naming) of every module on the hot path, and
the deterministic synthetic-weight generator used when no checkpoint is available (BASELINE.json: random-init SD /
CLIP / ODISE weights).  Names follow
  * ldm UNetModel / AutoencoderKL       -> `model.diffusion_model.*`, `first_stage_model.*`  
  * FeatureExtractorBackbone            -> `backbone.feature_projections.*`  (feature_extractor.py:53-66)
  * LdmImplicitCaptionerExtractor       -> `backbone.feature_extractor.*`    (ldm.py:651-670)
  * MSDeformAttnPixelDecoder            -> `sem_seg_head.pixel_decoder.*`    (msdeformattn.py:165-312)
  * ODISEMultiScaleMaskedTransformerDecoder -> `sem_seg_head.predictor.*`    (mask2former_transformer_decoder.py:236-340, odise.py:966-982)
  * CategoryEmbed                       -> `category_head.*`                 (odise.py:1219-1245)
so a real ODISE / SD checkpoint can be loaded through the same dict.
"""
import math

import torch

UNET_PREFIX = "model.diffusion_model."
VAE_PREFIX = "first_stage_model."


# ------------------------------------------------------------------------------------------------ UNet structure
def unet_blocks(model_channels=320, channel_mult=(1, 2, 4, 4), num_res_blocks=2, attention_resolutions=(4, 2, 1)):
    """Mirror of ldm UNetModel.__init__: returns (input_blocks, middle_block, output_blocks); each block is a list
    of ("conv_in", cin, cout) | ("res", cin, cout) | ("st", ch) | ("down", ch) | ("up", ch)."""
    mc = model_channels
    inp = [[("conv_in", 4, mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            layers = [("res", ch, mult * mc)]
            ch = mult * mc
            if ds in attention_resolutions:
                layers.append(("st", ch))
            inp.append(layers)
            chans.append(ch)
        if level != len(channel_mult) - 1:
            inp.append([("down", ch)])
            chans.append(ch)
            ds *= 2
    mid = [("res", ch, ch), ("st", ch), ("res", ch, ch)]
    out = []
    for level, mult in list(enumerate(channel_mult))[::-1]:
        for i in range(num_res_blocks + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, mc * mult)]
            ch = mc * mult
            if ds in attention_resolutions:
                layers.append(("st", ch))
            if level and i == num_res_blocks:
                layers.append(("up", ch))
                ds //= 2
            out.append(layers)
    return inp, mid, out


def _res_params(p, cin, cout, emb=1280):
    ps = [(p + "in_layers.0.weight", (cin,), "gamma"), (p + "in_layers.0.bias", (cin,), "beta"),
          (p + "in_layers.2.weight", (cout, cin, 3, 3), "w"), (p + "in_layers.2.bias", (cout,), "b"),
          (p + "emb_layers.1.weight", (cout, emb), "w"), (p + "emb_layers.1.bias", (cout,), "b"),
          (p + "out_layers.0.weight", (cout,), "gamma"), (p + "out_layers.0.bias", (cout,), "beta"),
          (p + "out_layers.3.weight", (cout, cout, 3, 3), "w"), (p + "out_layers.3.bias", (cout,), "b")]
    if cin != cout:
        ps += [(p + "skip_connection.weight", (cout, cin, 1, 1), "w"), (p + "skip_connection.bias", (cout,), "b")]
    return ps


def _st_params(p, ch, ctx=768):
    t = p + "transformer_blocks.0."
    ps = [(p + "norm.weight", (ch,), "gamma"), (p + "norm.bias", (ch,), "beta"),
          (p + "proj_in.weight", (ch, ch, 1, 1), "w"), (p + "proj_in.bias", (ch,), "b")]
    for a, kd in (("attn1", ch), ("attn2", ctx)):
        ps += [(t + a + ".to_q.weight", (ch, ch), "w"), (t + a + ".to_k.weight", (ch, kd), "w"),
               (t + a + ".to_v.weight", (ch, kd), "w"), (t + a + ".to_out.0.weight", (ch, ch), "w"),
               (t + a + ".to_out.0.bias", (ch,), "b")]
    ps += [(t + "ff.net.0.proj.weight", (8 * ch, ch), "w"), (t + "ff.net.0.proj.bias", (8 * ch,), "b"),
           (t + "ff.net.2.weight", (ch, 4 * ch), "w"), (t + "ff.net.2.bias", (ch,), "b")]
    for n in ("norm1", "norm2", "norm3"):
        ps += [(t + n + ".weight", (ch,), "gamma"), (t + n + ".bias", (ch,), "beta")]
    ps += [(p + "proj_out.weight", (ch, ch, 1, 1), "w"), (p + "proj_out.bias", (ch,), "b")]
    return ps


def _block_params(p, layers):
    ps = []
    for j, l in enumerate(layers):
        q = f"{p}{j}."
        if l[0] == "conv_in":
            ps += [(q + "weight", (l[2], l[1], 3, 3), "w"), (q + "bias", (l[2],), "b")]
        elif l[0] == "res":
            ps += _res_params(q, l[1], l[2])
        elif l[0] == "st":
            ps += _st_params(q, l[1])
        elif l[0] == "down":
            ps += [(q + "op.weight", (l[1], l[1], 3, 3), "w"), (q + "op.bias", (l[1],), "b")]
        elif l[0] == "up":
            ps += [(q + "conv.weight", (l[1], l[1], 3, 3), "w"), (q + "conv.bias", (l[1],), "b")]
    return ps


def unet_params(prefix=UNET_PREFIX):
    inp, mid, out = unet_blocks()
    ps = [(prefix + "time_embed.0.weight", (1280, 320), "w"), (prefix + "time_embed.0.bias", (1280,), "b"),
          (prefix + "time_embed.2.weight", (1280, 1280), "w"), (prefix + "time_embed.2.bias", (1280,), "b")]
    for i, layers in enumerate(inp):
        ps += _block_params(f"{prefix}input_blocks.{i}.", layers)
    ps += _block_params(prefix + "middle_block.", mid)
    for i, layers in enumerate(out):
        ps += _block_params(f"{prefix}output_blocks.{i}.", layers)
    ps += [(prefix + "out.0.weight", (320,), "gamma"), (prefix + "out.0.bias", (320,), "beta"),
           (prefix + "out.2.weight", (4, 320, 3, 3), "w"), (prefix + "out.2.bias", (4,), "b")]
    return ps


# ------------------------------------------------------------------------------------------------ VAE structure
def _vae_res(p, cin, cout):
    ps = [(p + "norm1.weight", (cin,), "gamma"), (p + "norm1.bias", (cin,), "beta"),
          (p + "conv1.weight", (cout, cin, 3, 3), "w"), (p + "conv1.bias", (cout,), "b"),
          (p + "norm2.weight", (cout,), "gamma"), (p + "norm2.bias", (cout,), "beta"),
          (p + "conv2.weight", (cout, cout, 3, 3), "w"), (p + "conv2.bias", (cout,), "b")]
    if cin != cout:
        ps += [(p + "nin_shortcut.weight", (cout, cin, 1, 1), "w"), (p + "nin_shortcut.bias", (cout,), "b")]
    return ps


def _vae_attn(p, c):
    ps = [(p + "norm.weight", (c,), "gamma"), (p + "norm.bias", (c,), "beta")]
    for n in ("q", "k", "v", "proj_out"):
        ps += [(p + n + ".weight", (c, c, 1, 1), "w"), (p + n + ".bias", (c,), "b")]
    return ps


def vae_params(prefix=VAE_PREFIX, ch=128, ch_mult=(1, 2, 4, 4), nres=2, z=4):
    e, d = prefix + "encoder.", prefix + "decoder."
    ps = [(e + "conv_in.weight", (ch, 3, 3, 3), "w"), (e + "conv_in.bias", (ch,), "b")]
    in_mult = (1,) + tuple(ch_mult)
    bi = ch
    for i in range(len(ch_mult)):
        bi, bo = ch * in_mult[i], ch * ch_mult[i]
        for j in range(nres):
            ps += _vae_res(f"{e}down.{i}.block.{j}.", bi, bo)
            bi = bo
        if i != len(ch_mult) - 1:
            ps += [(f"{e}down.{i}.downsample.conv.weight", (bi, bi, 3, 3), "w"),
                   (f"{e}down.{i}.downsample.conv.bias", (bi,), "b")]
    ps += _vae_res(e + "mid.block_1.", bi, bi) + _vae_attn(e + "mid.attn_1.", bi) + _vae_res(e + "mid.block_2.", bi, bi)
    ps += [(e + "norm_out.weight", (bi,), "gamma"), (e + "norm_out.bias", (bi,), "beta"),
           (e + "conv_out.weight", (2 * z, bi, 3, 3), "w"), (e + "conv_out.bias", (2 * z,), "b")]
    bi = ch * ch_mult[-1]
    ps += [(d + "conv_in.weight", (bi, z, 3, 3), "w"), (d + "conv_in.bias", (bi,), "b")]
    ps += _vae_res(d + "mid.block_1.", bi, bi) + _vae_attn(d + "mid.attn_1.", bi) + _vae_res(d + "mid.block_2.", bi, bi)
    for i in reversed(range(len(ch_mult))):
        bo = ch * ch_mult[i]
        for j in range(nres + 1):
            ps += _vae_res(f"{d}up.{i}.block.{j}.", bi, bo)
            bi = bo
        if i != 0:
            ps += [(f"{d}up.{i}.upsample.conv.weight", (bi, bi, 3, 3), "w"), (f"{d}up.{i}.upsample.conv.bias", (bi,), "b")]
    ps += [(d + "norm_out.weight", (bi,), "gamma"), (d + "norm_out.bias", (bi,), "beta"),
           (d + "conv_out.weight", (3, bi, 3, 3), "w"), (d + "conv_out.bias", (3,), "b")]
    ps += [(prefix + "quant_conv.weight", (2 * z, 2 * z, 1, 1), "w"), (prefix + "quant_conv.bias", (2 * z,), "b"),
           (prefix + "post_quant_conv.weight", (z, z, 1, 1), "w"), (prefix + "post_quant_conv.bias", (z,), "b")]
    return ps


# ------------------------------------------------------------------------------------------------ ODISE trainables
FEATURE_DIMS = (512, 512, 2560, 1920, 960, 640, 512, 512)   # enc(5,7) | unet(2,5,8,11) | dec(2,5)  (ldm.py:284-346)
FEATURE_STRIDES = (4, 8, 32, 32, 16, 8, 8, 4)               # clamped to [4, 32] (feature_extractor.py:88-99)


def backbone_params(prefix="backbone."):
    ps = []
    for i, cin in enumerate(FEATURE_DIMS):   # d2 BottleneckBlock(in, 128, 512, norm="GN")
        q = f"{prefix}feature_projections.{i}.0."
        if cin != 512:
            ps += [(q + "shortcut.weight", (512, cin, 1, 1), "w"), (q + "shortcut.norm.weight", (512,), "gamma"),
                   (q + "shortcut.norm.bias", (512,), "beta")]
        ps += [(q + "conv1.weight", (128, cin, 1, 1), "w"), (q + "conv1.norm.weight", (128,), "gamma"),
               (q + "conv1.norm.bias", (128,), "beta"),
               (q + "conv2.weight", (128, 128, 3, 3), "w"), (q + "conv2.norm.weight", (128,), "gamma"),
               (q + "conv2.norm.bias", (128,), "beta"),
               (q + "conv3.weight", (512, 128, 1, 1), "w"), (q + "conv3.norm.weight", (512,), "gamma"),
               (q + "conv3.norm.bias", (512,), "beta")]
    f = prefix + "feature_extractor."
    ps += [(f + "clip_project.linear.weight", (768, 768), "w"), (f + "clip_project.linear.bias", (768,), "b"),
           (f + "clip_project.positional_embedding", (1, 77, 768), "pos"),
           (f + "alpha_cond", (1, 77, 768), "alpha"),
           (f + "time_embed_project.linear.weight", (1280, 768), "w"), (f + "time_embed_project.linear.bias", (1280,), "b"),
           (f + "time_embed_project.positional_embedding", (1, 1, 1280), "pos"),
           (f + "alpha_cond_time_embed", (1280,), "alpha")]
    return ps


def pixel_decoder_params(prefix="sem_seg_head.pixel_decoder.", n_layers=6, C=256, ffn=1024, M=8, L=3, P=4):
    ps = []
    for i in range(3):
        ps += [(f"{prefix}input_proj.{i}.0.weight", (C, 512, 1, 1), "w"), (f"{prefix}input_proj.{i}.0.bias", (C,), "b"),
               (f"{prefix}input_proj.{i}.1.weight", (C,), "gamma"), (f"{prefix}input_proj.{i}.1.bias", (C,), "beta")]
    ps += [(prefix + "transformer.level_embed", (L, C), "emb")]
    for l in range(n_layers):
        q = f"{prefix}transformer.encoder.layers.{l}."
        ps += [(q + "self_attn.sampling_offsets.weight", (M * L * P * 2, C), "msda_off_w"),
               (q + "self_attn.sampling_offsets.bias", (M * L * P * 2,), "msda_off_b"),
               (q + "self_attn.attention_weights.weight", (M * L * P, C), "w"),
               (q + "self_attn.attention_weights.bias", (M * L * P,), "b"),
               (q + "self_attn.value_proj.weight", (C, C), "w"), (q + "self_attn.value_proj.bias", (C,), "b"),
               (q + "self_attn.output_proj.weight", (C, C), "w"), (q + "self_attn.output_proj.bias", (C,), "b"),
               (q + "norm1.weight", (C,), "gamma"), (q + "norm1.bias", (C,), "beta"),
               (q + "linear1.weight", (ffn, C), "w"), (q + "linear1.bias", (ffn,), "b"),
               (q + "linear2.weight", (C, ffn), "w"), (q + "linear2.bias", (C,), "b"),
               (q + "norm2.weight", (C,), "gamma"), (q + "norm2.bias", (C,), "beta")]
    ps += [(prefix + "mask_features.weight", (C, C, 1, 1), "w"), (prefix + "mask_features.bias", (C,), "b"),
           (prefix + "adapter_1.weight", (C, 512, 1, 1), "w"), (prefix + "adapter_1.norm.weight", (C,), "gamma"),
           (prefix + "adapter_1.norm.bias", (C,), "beta"),
           (prefix + "layer_1.weight", (C, C, 3, 3), "w"), (prefix + "layer_1.norm.weight", (C,), "gamma"),
           (prefix + "layer_1.norm.bias", (C,), "beta")]
    return ps


def decoder_params(prefix="sem_seg_head.predictor.", n_layers=9, C=256, ffn=2048, Q=100, n_levels=3):
    ps = []
    for i in range(n_layers):
        for nm, attn in ((f"transformer_self_attention_layers.{i}", "self_attn"),
                         (f"transformer_cross_attention_layers.{i}", "multihead_attn")):
            q = f"{prefix}{nm}."
            ps += [(q + attn + ".in_proj_weight", (3 * C, C), "w"), (q + attn + ".in_proj_bias", (3 * C,), "b"),
                   (q + attn + ".out_proj.weight", (C, C), "w"), (q + attn + ".out_proj.bias", (C,), "b"),
                   (q + "norm.weight", (C,), "gamma"), (q + "norm.bias", (C,), "beta")]
        q = f"{prefix}transformer_ffn_layers.{i}."
        ps += [(q + "linear1.weight", (ffn, C), "w"), (q + "linear1.bias", (ffn,), "b"),
               (q + "linear2.weight", (C, ffn), "w"), (q + "linear2.bias", (C,), "b"),
               (q + "norm.weight", (C,), "gamma"), (q + "norm.bias", (C,), "beta")]
    ps += [(prefix + "decoder_norm.weight", (C,), "gamma"), (prefix + "decoder_norm.bias", (C,), "beta"),
           (prefix + "query_feat.weight", (Q, C), "emb"), (prefix + "query_embed.weight", (Q, C), "emb"),
           (prefix + "level_embed.weight", (n_levels, C), "emb")]
    for base in ("mask_embed.", "post_mask_embed.mask_embed.1."):
        for j in range(3):
            ps += [(f"{prefix}{base}layers.{j}.weight", (C, C), "w"), (f"{prefix}{base}layers.{j}.bias", (C,), "b")]
    ps += [(prefix + "post_mask_embed.pool_proj.0.weight", (C,), "gamma"), (prefix + "post_mask_embed.pool_proj.0.bias", (C,), "beta"),
           (prefix + "post_mask_embed.pool_proj.1.weight", (C, C), "w"), (prefix + "post_mask_embed.pool_proj.1.bias", (C,), "b"),
           (prefix + "post_mask_embed.mask_embed.0.weight", (C,), "gamma"), (prefix + "post_mask_embed.mask_embed.0.bias", (C,), "beta"),
           (prefix + "post_mask_embed.logit_scale", (), "logit_scale")]
    return ps


def category_head_params(prefix="category_head."):
    return [(prefix + "text_proj.weight", (256, 768), "w"), (prefix + "text_proj.bias", (256,), "b")]


CLIP_PREFIX = "clip.visual."


def clip_visual_params(prefix=CLIP_PREFIX, width=1024, layers=24, patch=14, image=336, out_dim=768):
    """open_clip VisionTransformer (ViT-L-14-336, pretrained "openai") parameter names — the image tower behind
    ClipAdapter.embed_image (odise/modeling/meta_arch/clip.py:177-231): conv1 (no bias), class / positional embedding,
    ln_pre, `layers` ResidualAttentionBlocks (ln_1, attn.in_proj, attn.out_proj, ln_2, mlp.c_fc, mlp.c_proj), ln_post, proj."""
    n_tok = (image // patch) ** 2 + 1
    ps = [(prefix + "conv1.weight", (width, 3, patch, patch), "w"), (prefix + "class_embedding", (width,), "cls"),
          (prefix + "positional_embedding", (n_tok, width), "cls"),
          (prefix + "ln_pre.weight", (width,), "gamma"), (prefix + "ln_pre.bias", (width,), "beta")]
    for i in range(layers):
        q = f"{prefix}transformer.resblocks.{i}."
        ps += [(q + "ln_1.weight", (width,), "gamma"), (q + "ln_1.bias", (width,), "beta"),
               (q + "attn.in_proj_weight", (3 * width, width), "w"), (q + "attn.in_proj_bias", (3 * width,), "b"),
               (q + "attn.out_proj.weight", (width, width), "w"), (q + "attn.out_proj.bias", (width,), "b"),
               (q + "ln_2.weight", (width,), "gamma"), (q + "ln_2.bias", (width,), "beta"),
               (q + "mlp.c_fc.weight", (4 * width, width), "w"), (q + "mlp.c_fc.bias", (4 * width,), "b"),
               (q + "mlp.c_proj.weight", (width, 4 * width), "w"), (q + "mlp.c_proj.bias", (width,), "b")]
    ps += [(prefix + "ln_post.weight", (width,), "gamma"), (prefix + "ln_post.bias", (width,), "beta"),
           (prefix + "proj", (width, out_dim), "proj")]
    return ps


CLIP_TEXT_PREFIX = "clip."
SD_TEXT_PREFIX = "cond_stage_model.transformer.text_model."


def clip_text_params(prefix=CLIP_TEXT_PREFIX, width=768, layers=12, vocab=49408, ctx=77, out_dim=768):
    """open_clip CLIP text tower (ViT-L-14-336 "openai"), OpenAI parameter names — what build_clip_text_embed /
    ClipAdapter._encode_text run (odise/modeling/meta_arch/clip.py:29-73, :138-152)."""
    ps = [(prefix + "token_embedding.weight", (vocab, width), "pos"), (prefix + "positional_embedding", (ctx, width), "pos")]
    for i in range(layers):
        q = f"{prefix}transformer.resblocks.{i}."
        ps += [(q + "ln_1.weight", (width,), "gamma"), (q + "ln_1.bias", (width,), "beta"),
               (q + "attn.in_proj_weight", (3 * width, width), "w"), (q + "attn.in_proj_bias", (3 * width,), "b"),
               (q + "attn.out_proj.weight", (width, width), "w"), (q + "attn.out_proj.bias", (width,), "b"),
               (q + "ln_2.weight", (width,), "gamma"), (q + "ln_2.bias", (width,), "beta"),
               (q + "mlp.c_fc.weight", (4 * width, width), "w"), (q + "mlp.c_fc.bias", (4 * width,), "b"),
               (q + "mlp.c_proj.weight", (width, 4 * width), "w"), (q + "mlp.c_proj.bias", (width,), "b")]
    ps += [(prefix + "ln_final.weight", (width,), "gamma"), (prefix + "ln_final.bias", (width,), "beta"),
           (prefix + "text_projection", (width, out_dim), "proj"), (prefix + "logit_scale", (), "logit_scale")]
    return ps


def sd_text_params(prefix=SD_TEXT_PREFIX, width=768, layers=12, vocab=49408, ctx=77):
    """SD-v1 `cond_stage_model` (ldm FrozenCLIPEmbedder = HF CLIPTextModel openai/clip-vit-large-patch14), HF parameter
    names as stored in sd-v1-*.ckpt; produces `uncond_inputs` = last_hidden_state of "" (ldm.py:116)."""
    ps = [(prefix + "embeddings.token_embedding.weight", (vocab, width), "pos"),
          (prefix + "embeddings.position_embedding.weight", (ctx, width), "pos")]
    for i in range(layers):
        q = f"{prefix}encoder.layers.{i}."
        ps += [(q + "layer_norm1.weight", (width,), "gamma"), (q + "layer_norm1.bias", (width,), "beta")]
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            ps += [(q + f"self_attn.{n}.weight", (width, width), "w"), (q + f"self_attn.{n}.bias", (width,), "b")]
        ps += [(q + "layer_norm2.weight", (width,), "gamma"), (q + "layer_norm2.bias", (width,), "beta"),
               (q + "mlp.fc1.weight", (4 * width, width), "w"), (q + "mlp.fc1.bias", (4 * width,), "b"),
               (q + "mlp.fc2.weight", (width, 4 * width), "w"), (q + "mlp.fc2.bias", (width,), "b")]
    ps += [(prefix + "final_layer_norm.weight", (width,), "gamma"), (prefix + "final_layer_norm.bias", (width,), "beta")]
    return ps


def hf_text_to_openai(sd, src_prefix=SD_TEXT_PREFIX, dst_prefix="sd_text."):
    """HF CLIPTextModel names -> OpenAI CLIP names (q/k/v concatenated into in_proj), so one engine serves both."""
    out = {dst_prefix + "token_embedding.weight": sd[src_prefix + "embeddings.token_embedding.weight"],
           dst_prefix + "positional_embedding": sd[src_prefix + "embeddings.position_embedding.weight"],
           dst_prefix + "ln_final.weight": sd[src_prefix + "final_layer_norm.weight"],
           dst_prefix + "ln_final.bias": sd[src_prefix + "final_layer_norm.bias"]}
    i = 0
    while f"{src_prefix}encoder.layers.{i}.layer_norm1.weight" in sd:
        s, d = f"{src_prefix}encoder.layers.{i}.", f"{dst_prefix}transformer.resblocks.{i}."
        for a, b in (("layer_norm1", "ln_1"), ("layer_norm2", "ln_2"), ("self_attn.out_proj", "attn.out_proj"),
                     ("mlp.fc1", "mlp.c_fc"), ("mlp.fc2", "mlp.c_proj")):
            out[d + b + ".weight"], out[d + b + ".bias"] = sd[s + a + ".weight"], sd[s + a + ".bias"]
        out[d + "attn.in_proj_weight"] = torch.cat([sd[s + f"self_attn.{n}.weight"] for n in ("q_proj", "k_proj", "v_proj")])
        out[d + "attn.in_proj_bias"] = torch.cat([sd[s + f"self_attn.{n}.bias"] for n in ("q_proj", "k_proj", "v_proj")])
        i += 1
    return out


def msda_offset_bias(M=8, L=3, P=4):
    """MSDeformAttn._reset_parameters directional grid (ops/modules/ms_deform_attn.py:66-74)."""
    thetas = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
    grid = torch.stack([thetas.cos(), thetas.sin()], -1)
    grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(M, 1, 1, 2).repeat(1, L, P, 1)
    for i in range(P):
        grid[:, :, i, :] *= i + 1
    return grid.reshape(-1)


def synth_state_dict(params, seed=0, dtype=torch.float32):
    """Deterministic synthetic weights: conv / linear weights N(0, 1/fan_in) (variance preserving,
    every ldm / M2F zero-init overwritten), biases N(0, 0.02^2), norm gamma 1 + N(0, 0.1^2), beta N(0, 0.1^2),
    embeddings N(0, 1), MSDeformAttn offsets N(0, 0.05^2) on top of the reference's directional bias,
    alpha_* N(0, 0.5^2), logit_scale ln(1/0.07).  One CPU generator, parameters drawn in list order."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape, kind in params:
        if kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
        elif kind == "b":
            t = torch.randn(shape, generator=g) * 0.02
        elif kind == "gamma":
            t = 1.0 + torch.randn(shape, generator=g) * 0.1
        elif kind == "beta":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "emb":
            t = torch.randn(shape, generator=g)
        elif kind == "pos":
            t = torch.randn(shape, generator=g) * 0.02
        elif kind == "alpha":
            t = torch.randn(shape, generator=g) * 0.5
        elif kind == "msda_off_w":
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == "msda_off_b":
            t = msda_offset_bias()
        elif kind == "logit_scale":
            t = torch.tensor(math.log(1 / 0.07))
        elif kind == "cls":          # CLIP class / positional embeddings: scale = width ** -0.5 like open_clip
            t = torch.randn(shape, generator=g) * (shape[-1] ** -0.5)
        elif kind == "proj":         # [width, out]: used as x @ proj
            t = torch.randn(shape, generator=g) * (shape[0] ** -0.5)
        else:
            raise ValueError(kind)
        sd[name] = t.to(dtype)
    return sd


def head_params():
    return pixel_decoder_params() + decoder_params() + category_head_params()
