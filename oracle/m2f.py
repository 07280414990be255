"""Oracle: Mask2Former pixel decoder + ODISE masked-attention decoder + open-vocabulary scoring, and the
FeatureExtractorBackbone projections (CPU, functional PyTorch over a state dict with the REFERENCE's parameter
names).  TEST INFRASTRUCTURE ONLY.

PINNED: tests/test_oracle_cpu.py loads the same state dict into the reference's own modules (imported verbatim from
/root/reference through oracle/refshim.py) and requires these functions to reproduce them; tools/make_golden_head.py
stores the reference outputs in tests/golden/head_*.pt for the machines that have no /root/reference.
Only d2 BottleneckBlock (detectron2 v0.6, not vendored) is restated from the published definition ("unpinned").

M2F = third_party/Mask2Former/mask2former/modeling.
"""
import math

import torch
import torch.nn.functional as F

from .msda import msda_forward


def _get(sd, prefix):
    return sd[prefix + ".weight"], sd.get(prefix + ".bias")


def linear(sd, prefix, x):
    w, b = _get(sd, prefix)
    return F.linear(x, w, b)


def layer_norm(sd, prefix, x):
    w, b = _get(sd, prefix)
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def group_norm(sd, prefix, x, groups=32, eps=1e-5):
    w, b = _get(sd, prefix)
    return F.group_norm(x, groups, w, b, eps)


def position_embedding_sine(B, H, W, num_pos_feats=128, temperature=10000, dtype=torch.float32):
    """PositionEmbeddingSine(normalize=True) with an all-False mask (M2F/transformer_decoder/position_encoding.py:29-52)."""
    scale, eps = 2 * math.pi, 1e-6
    y_embed = torch.arange(1, H + 1, dtype=torch.float32).view(1, H, 1).expand(B, H, W)
    x_embed = torch.arange(1, W + 1, dtype=torch.float32).view(1, 1, W).expand(B, H, W)
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2).to(dtype)


# ----------------------------------------------------------------------------------------------- pixel decoder
def reference_points(spatial_shapes, B, dtype=torch.float32):
    """MSDeformAttnTransformerEncoder.get_reference_points with valid_ratios == 1 (M2F/pixel_decoder/msdeformattn.py:141-153)."""
    pts = []
    for H, W in spatial_shapes:
        ry, rx = torch.meshgrid(torch.linspace(0.5, H - 0.5, H, dtype=torch.float32),
                                torch.linspace(0.5, W - 0.5, W, dtype=torch.float32), indexing="ij")
        pts.append(torch.stack((rx.reshape(-1) / W, ry.reshape(-1) / H), -1))
    ref = torch.cat(pts, 0)[None].expand(B, -1, -1)
    L = len(spatial_shapes)
    return ref[:, :, None, :].expand(-1, -1, L, -1).to(dtype)


def msdeform_attn(sd, prefix, query, ref_points, src, spatial_shapes, level_start, M=8, P=4):
    """MSDeformAttn.forward (M2F/pixel_decoder/ops/modules/ms_deform_attn.py:82-125)."""
    N, Lq, C = query.shape
    L = spatial_shapes.shape[0]
    value = linear(sd, prefix + ".value_proj", src).view(N, -1, M, C // M)
    off = linear(sd, prefix + ".sampling_offsets", query).view(N, Lq, M, L, P, 2)
    aw = linear(sd, prefix + ".attention_weights", query).view(N, Lq, M, L * P)
    aw = F.softmax(aw, -1).view(N, Lq, M, L, P)
    normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1).to(off.dtype)
    loc = ref_points[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
    out = msda_forward(value, spatial_shapes, level_start, loc, aw)
    return linear(sd, prefix + ".output_proj", out)


def pixel_decoder(sd, features, prefix="", n_layers=6):
    """MSDeformAttnPixelDecoder.forward_features (M2F/pixel_decoder/msdeformattn.py:314-358) for
    transformer_in_features = [s3, s4, s5], one FPN level (s2), common_stride 4.
    features: dict s2..s5 of NCHW tensors.  Returns (mask_features, out[0], multi_scale_features)."""
    p = prefix
    srcs, pos = [], []
    for idx, f in enumerate(["s5", "s4", "s3"]):
        x = features[f].float()
        w, b = _get(sd, f"{p}input_proj.{idx}.0")
        srcs.append(group_norm(sd, f"{p}input_proj.{idx}.1", F.conv2d(x, w, b)))
        pos.append(position_embedding_sine(x.shape[0], x.shape[2], x.shape[3], dtype=x.dtype))
    B = srcs[0].shape[0]
    shapes = [(s.shape[2], s.shape[3]) for s in srcs]
    lvl = sd[f"{p}transformer.level_embed"]
    src_flat = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    pos_flat = torch.cat([pe.flatten(2).transpose(1, 2) + lvl[i].view(1, 1, -1) for i, pe in enumerate(pos)], 1)
    ss = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
    ref = reference_points(shapes, B, src_flat.dtype)
    out = src_flat
    for l in range(n_layers):   # MSDeformAttnTransformerEncoderLayer.forward (:122-131), dropout 0
        lp = f"{p}transformer.encoder.layers.{l}"
        src2 = msdeform_attn(sd, lp + ".self_attn", out + pos_flat, ref, out, ss, lsi)
        out = layer_norm(sd, lp + ".norm1", out + src2)
        src2 = linear(sd, lp + ".linear2", F.relu(linear(sd, lp + ".linear1", out)))
        out = layer_norm(sd, lp + ".norm2", out + src2)
    ys = torch.split(out, [h * w for h, w in shapes], dim=1)
    outs = [z.transpose(1, 2).reshape(B, -1, h, w) for z, (h, w) in zip(ys, shapes)]
    # FPN level on s2 (:343-351); d2 Conv2d(norm=GN) keeps GN params under "<name>.norm"
    x = features["s2"].float()
    cur = group_norm(sd, f"{p}adapter_1.norm", F.conv2d(x, sd[f"{p}adapter_1.weight"]))
    y = cur + F.interpolate(outs[-1], size=cur.shape[-2:], mode="bilinear", align_corners=False)
    y = F.relu(group_norm(sd, f"{p}layer_1.norm", F.conv2d(y, sd[f"{p}layer_1.weight"], padding=1)))
    outs.append(y)
    w, b = _get(sd, f"{p}mask_features")
    return F.conv2d(outs[-1], w, b), outs[0], outs[:3]


# ----------------------------------------------------------------------------------------------- decoder
def mha(sd, prefix, query, key, value, attn_mask=None, heads=8):
    """nn.MultiheadAttention forward, seq-first [T, B, C], dropout 0 (torch F.multi_head_attention_forward)."""
    Tq, B, C = query.shape
    Tk = key.shape[0]
    w, b = sd[prefix + ".in_proj_weight"], sd[prefix + ".in_proj_bias"]
    q = F.linear(query, w[:C], b[:C])
    k = F.linear(key, w[C:2 * C], b[C:2 * C])
    v = F.linear(value, w[2 * C:], b[2 * C:])
    d = C // heads
    q = q.reshape(Tq, B * heads, d).transpose(0, 1) * (d ** -0.5)
    k = k.reshape(Tk, B * heads, d).transpose(0, 1)
    v = v.reshape(Tk, B * heads, d).transpose(0, 1)
    s = torch.bmm(q, k.transpose(1, 2))
    if attn_mask is not None:
        s = s.masked_fill(attn_mask, float("-inf"))
    o = torch.bmm(F.softmax(s, dim=-1), v).transpose(0, 1).reshape(Tq, B, C)
    return linear(sd, prefix + ".out_proj", o)


def mlp3(sd, prefix, x):
    """M2F MLP(num_layers=3) (mask2former_transformer_decoder.py:192-205)."""
    x = F.relu(linear(sd, prefix + ".layers.0", x))
    x = F.relu(linear(sd, prefix + ".layers.1", x))
    return linear(sd, prefix + ".layers.2", x)


def mask_pooling(x, mask):
    """MaskPooling.forward, hard pooling (odise.py:937-963)."""
    m = (mask.sigmoid() > 0.5).to(mask.dtype)
    denorm = m.sum(dim=(-1, -2), keepdim=True) + 1e-8
    return torch.einsum("bchw,bqhw->bqc", x, m / denorm)


def prediction_heads(sd, p, output, mask_features, target_size, heads=8, forced_masks=None):
    """ODISEMultiScaleMaskedTransformerDecoder.forward_prediction_heads + PooledMaskEmbed.forward
    (odise.py:729-776, :984-1015).  forced_masks: teacher-forced outputs_mask for discontinuity-free tests."""
    dec = layer_norm(sd, p + "decoder_norm", output).transpose(0, 1)
    mask_embed = mlp3(sd, p + "mask_embed", dec)
    outputs_mask = torch.einsum("bqc,bchw->bqhw", mask_embed, mask_features)
    masks_for_threshold = outputs_mask if forced_masks is None else forced_masks
    pooled = mask_pooling(mask_features, masks_for_threshold)
    pooled = linear(sd, p + "post_mask_embed.pool_proj.1", layer_norm(sd, p + "post_mask_embed.pool_proj.0", pooled))
    pooled = pooled + dec
    me = mlp3(sd, p + "post_mask_embed.mask_embed.1", layer_norm(sd, p + "post_mask_embed.mask_embed.0", pooled))
    logit_scale = torch.clamp(sd[p + "post_mask_embed.logit_scale"].exp(), max=100)
    attn_mask = F.interpolate(masks_for_threshold, size=target_size, mode="bilinear", align_corners=False)
    attn_mask = (attn_mask.sigmoid().flatten(2).unsqueeze(1).repeat(1, heads, 1, 1).flatten(0, 1) < 0.5).bool()
    return outputs_mask, attn_mask, dict(mask_embed=me, mask_pooled_features=pooled, logit_scale=logit_scale)


def transformer_decoder(sd, x, mask_features, prefix="", n_layers=9, heads=8, forced_masks=None):
    """ODISEMultiScaleMaskedTransformerDecoder.forward (odise.py:642-727).  x: 3 multi-scale NCHW maps.
    forced_masks: optional list of n_layers+1 mask-logit tensors used for every threshold (attention masks and
    hard pooling) instead of the self-computed ones.  Returns dict + per-head list of pre-threshold mask logits."""
    p = prefix
    src, pos, sizes = [], [], []
    for i in range(3):
        sizes.append(x[i].shape[-2:])
        pe = position_embedding_sine(x[i].shape[0], x[i].shape[2], x[i].shape[3], dtype=x[i].dtype)
        pos.append(pe.flatten(2).permute(2, 0, 1))
        s = x[i].flatten(2) + sd[p + "level_embed.weight"][i][None, :, None]   # input_proj = identity (256 == 256)
        src.append(s.permute(2, 0, 1))
    B = src[0].shape[1]
    query_embed = sd[p + "query_embed.weight"].unsqueeze(1).repeat(1, B, 1)
    output = sd[p + "query_feat.weight"].unsqueeze(1).repeat(1, B, 1)
    all_masks, all_extra = [], []
    fm = (lambda i: None) if forced_masks is None else (lambda i: forced_masks[i])
    om, attn_mask, extra = prediction_heads(sd, p, output, mask_features, sizes[0], heads, fm(0))
    all_masks.append(om)
    all_extra.append(extra)
    for i in range(n_layers):
        lvl = i % 3
        attn_mask[torch.where(attn_mask.sum(-1) == attn_mask.shape[-1])] = False
        cp = f"{p}transformer_cross_attention_layers.{i}"
        tgt2 = mha(sd, cp + ".multihead_attn", output + query_embed, src[lvl] + pos[lvl], src[lvl], attn_mask, heads)
        output = layer_norm(sd, cp + ".norm", output + tgt2)
        sp = f"{p}transformer_self_attention_layers.{i}"
        qk = output + query_embed
        tgt2 = mha(sd, sp + ".self_attn", qk, qk, output, None, heads)
        output = layer_norm(sd, sp + ".norm", output + tgt2)
        fp = f"{p}transformer_ffn_layers.{i}"
        tgt2 = linear(sd, fp + ".linear2", F.relu(linear(sd, fp + ".linear1", output)))
        output = layer_norm(sd, fp + ".norm", output + tgt2)
        om, attn_mask, extra = prediction_heads(sd, p, output, mask_features, sizes[(i + 1) % 3], heads, fm(i + 1))
        all_masks.append(om)
        all_extra.append(extra)
    out = dict(pred_masks=all_masks[-1], **all_extra[-1])
    out["aux_outputs"] = [dict(pred_masks=m, **e) for m, e in zip(all_masks[:-1], all_extra[:-1])]
    return out, all_masks


# ----------------------------------------------------------------------------------------------- scoring
def cal_pred_logits(mask_embed, text_embed, null_embed, logit_scale, group_sizes):
    """CategoryODISE.cal_pred_logits + ensemble_logits_with_labels("max") (odise.py:181-207, helper.py:79-109).
    group_sizes[i] = number of prompt columns of class i (len(labels[i]))."""
    me = F.normalize(mask_embed, dim=-1)
    te = F.normalize(text_embed, dim=-1)
    pred = logit_scale * (me @ te.t())
    outs, s = [], 0
    for n in group_sizes:
        outs.append(pred[..., s:s + n].max(dim=-1).values)
        s += n
    pred = torch.stack(outs, -1)
    null_pred = logit_scale * (me @ F.normalize(null_embed, dim=-1).t())
    return torch.cat([pred, null_pred], dim=-1)


def category_embed(sd, text_bank, null_bank, prefix="category_head."):
    """CategoryEmbed.forward eval branch (odise.py:1298-1307): text_proj on the cached CLIP text bank / null embed."""
    return linear(sd, prefix + "text_proj", text_bank), linear(sd, prefix + "text_proj", null_bank)


# ----------------------------------------------------------------------------------------------- backbone tail
def bottleneck_block(sd, prefix, x):
    """detectron2 v0.6 BottleneckBlock(norm="GN", stride 1): 1x1 -> GN -> ReLU -> 3x3 -> GN -> ReLU -> 1x1 -> GN,
    (+ 1x1 + GN shortcut iff in != out), ReLU; all convs bias-free.  [unpinned: d2 is not vendored]"""
    out = F.relu(group_norm(sd, prefix + ".conv1.norm", F.conv2d(x, sd[prefix + ".conv1.weight"])))
    out = F.relu(group_norm(sd, prefix + ".conv2.norm", F.conv2d(out, sd[prefix + ".conv2.weight"], padding=1)))
    out = group_norm(sd, prefix + ".conv3.norm", F.conv2d(out, sd[prefix + ".conv3.weight"]))
    if (prefix + ".shortcut.weight") in sd:
        x = group_norm(sd, prefix + ".shortcut.norm", F.conv2d(x, sd[prefix + ".shortcut.weight"]))
    return F.relu(out + x)


# (feature idx -> output stride) of FeatureExtractorBackbone for the ODISE config
# (feature_extractor.py:88-112 with feature_strides [4,8 | 64,32,16,8 | 8,4] clamped to [4,32])
FEATURE_STRIDES = [4, 8, 32, 32, 16, 8, 8, 4]
FEATURE_DIMS = [512, 512, 2560, 1920, 960, 640, 512, 512]


def forward_features(sd, feats, input_hw, prefix="backbone.feature_projections."):
    """FeatureExtractorBackbone.forward_features (feature_extractor.py:157-179): nearest-resize each tap to
    input/stride, BottleneckBlock projection, sum per stride -> {s2, s3, s4, s5}."""
    out = {}
    for idx, f in enumerate(feats):
        s = FEATURE_STRIDES[idx]
        r = F.interpolate(f, size=(input_hw[0] // s, input_hw[1] // s))
        y = bottleneck_block(sd, f"{prefix}{idx}.0", r)
        name = f"s{int(math.log2(s))}"
        out[name] = y if name not in out else out[name] + y
    return out
