"""Oracle: CLIP ViT image tower as used by ClipAdapter.embed_image (odise/modeling/meta_arch/clip.py:177-231).
TEST INFRASTRUCTURE ONLY.

The GLUE (_encode_image: conv1 -> tokens -> class/pos embedding -> ln_pre -> transformer -> ln_post -> proj -> token 0)
is vendored in the reference and PINNED: tests/test_oracle_cpu.py runs the reference's ClipAdapter._encode_image
verbatim on a fake `self.clip.visual` built from these modules.  The transformer block itself comes from
open-clip-torch==2.0.2 (setup.py:85), un-vendored: restated (PARITY UNPINNED) as
    x = x + attn(ln_1(x));  x = x + c_proj(QuickGELU(c_fc(ln_2(x))))          (ResidualAttentionBlock)
with nn.MultiheadAttention and QuickGELU(x) = x * sigmoid(1.702 x) (OpenAI-pretrained configs use quick_gelu).
Preprocessing: T.Resize(336, BICUBIC) + CenterCrop(336) + Normalize (clip.py:94; torchvision 0.14 = no antialias on
tensors — SURVEY.md §7).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d, heads):
        super().__init__()
        self.ln_1 = nn.LayerNorm(d)
        self.attn = nn.MultiheadAttention(d, heads)
        self.ln_2 = nn.LayerNorm(d)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(d, 4 * d))
        self.mlp.add_module("gelu", QuickGELU())
        self.mlp.add_module("c_proj", nn.Linear(4 * d, d))

    def forward(self, x, attn_mask=None):      # [L, N, D]; bool attn_mask: True = may NOT attend
        y = self.ln_1(x)
        x = x + self.attn(y, y, y, need_weights=False, attn_mask=attn_mask)[0]
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(width, heads) for _ in range(layers)])

    def forward(self, x, attn_mask=None):
        for r in self.resblocks:
            x = r(x, attn_mask)
        return x


class VisionTransformer(nn.Module):
    def __init__(self, image_size=336, patch=14, width=1024, layers=24, heads=16, out_dim=768):
        super().__init__()
        self.image_size = image_size
        self.conv1 = nn.Conv2d(3, width, patch, stride=patch, bias=False)
        self.class_embedding = nn.Parameter(torch.zeros(width))
        self.positional_embedding = nn.Parameter(torch.zeros((image_size // patch) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(torch.zeros(width, out_dim))


def preprocess(image, size=336):
    """clip_preprocess (clip.py:94): Resize(size, bicubic, no antialias) -> CenterCrop(size) -> Normalize."""
    B, C, H, W = image.shape
    if H <= W:
        nh, nw = size, int(size * W / H)
    else:
        nh, nw = int(size * H / W), size
    x = F.interpolate(image, size=(nh, nw), mode="bicubic", align_corners=False, antialias=False)
    top, left = int(round((nh - size) / 2.0)), int(round((nw - size) / 2.0))
    x = x[:, :, top:top + size, left:left + size]
    mean = torch.tensor(CLIP_MEAN, dtype=x.dtype).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=x.dtype).view(1, 3, 1, 1)
    return (x - mean) / std


def encode_image(visual, image):
    """ClipAdapter._encode_image (clip.py:177-222) -> image_embed [B, out_dim] (token 0 after ln_post + proj)."""
    x = visual.conv1(image)
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
    x = torch.cat([visual.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype), x], dim=1)
    x = x + visual.positional_embedding.to(x.dtype)
    x = visual.ln_pre(x)
    x = visual.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)
    x = visual.ln_post(x)
    return x[:, 0, :] @ visual.proj


def embed_image(visual, crop01):
    """ClipAdapter.embed_image with normalize=False (clip.py:225-231, ldm.py:652): crop in [0, 1] -> [B, out_dim]."""
    return encode_image(visual, preprocess(crop01, visual.image_size)).float()


# ---------------------------------------------------------------------------------------------------- MaskCLIP
# MaskCLIP (clip.py:239-361): Q extra "mask tokens" (copies of the class token after ln_pre) ride through the frozen
# ViT; nobody attends to them, and mask token q attends to the class token and to the patches its mask touches.
# Pinned: tests/test_oracle_cpu.py runs the reference's MaskCLIP methods verbatim on these modules.

def mask_clip_forward(visual, x, attn_mask, num_mask_tokens):
    """MaskCLIP._mask_clip_forward (clip.py:252-282): normalised image [B,3,S,S] -> [B, Q, out_dim]."""
    x = visual.conv1(x)
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
    x = torch.cat([visual.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype), x], dim=1)
    x = visual.ln_pre(x + visual.positional_embedding.to(x.dtype)).permute(1, 0, 2)
    x = torch.cat([x[0:1].expand(num_mask_tokens, -1, -1), x], dim=0)
    x = visual.transformer(x, attn_mask).permute(1, 0, 2)
    x = visual.ln_post(x[:, :num_mask_tokens, :])
    return torch.einsum("nld,dc->nlc", x, visual.proj)


def mask_attention_mask(visual, mask_logits):
    """encode_image_with_mask's mask construction (clip.py:291-321): mask logits [B, Q, S, S] (already at the CLIP
    resolution) -> bool [B*heads, Q+1+G*G, Q+1+G*G], True = blocked."""
    B, Q = mask_logits.shape[:2]
    patch_mask = F.max_pool2d(mask_logits.sigmoid(), kernel_size=visual.conv1.kernel_size, stride=visual.conv1.stride)
    blocked = (patch_mask < 0.5).reshape(B, Q, -1)
    n_img = visual.positional_embedding.shape[0] - 1
    n_all = Q + 1 + n_img
    am = torch.zeros(n_all, n_all, dtype=torch.bool)
    am[:, :Q] = True
    am = am.unsqueeze(0).repeat_interleave(B, dim=0)
    am[:, :Q, -n_img:] = blocked
    heads = visual.conv1.out_channels // 64
    return am.unsqueeze(1).expand(-1, heads, -1, -1).reshape(B * heads, n_all, n_all)


def get_mask_embed(visual, image01, mask_logits):
    """MaskCLIP.get_mask_embed (clip.py:325-339): image in [0,1] at any size, mask logits [B, Q, h, w] -> [B, Q, out]."""
    S = visual.image_size
    image = F.interpolate(image01, size=(S, S), mode="bilinear", align_corners=False)
    mask = F.interpolate(mask_logits, size=(S, S), mode="bilinear", align_corners=False)
    image = preprocess(image, S)                     # Resize(S) of an SxS tensor is the identity; Normalize remains
    return mask_clip_forward(visual, image, mask_attention_mask(visual, mask), mask_logits.shape[1])


def maskclip_pred_logits(mask_embed, text_embed, group_sizes, logit_scale):
    """MaskCLIP.pred_logits (clip.py:341-351) + ensemble_logits_with_labels (helper.py:79-109, "max").
    logit_scale = clamp(exp(clip.logit_scale), max=100) (clip.py:247-250)."""
    lg = torch.einsum("bqc,nc->bqn", F.normalize(mask_embed, dim=-1), F.normalize(text_embed, dim=-1)) * logit_scale
    out, o = [], 0
    for n in group_sizes:
        out.append(lg[..., o:o + n].max(dim=-1).values)
        o += n
    return torch.stack(out, dim=-1)


def pooling_clip_ensemble(pred_open_logits, mask_pred_open_logits, overlapping, alpha, beta):
    """PoolingCLIPHead.forward, normalize_logits=True branch (odise.py:1506-1536): geometric ensemble of the two class
    distributions with exponent alpha for classes seen in training (overlapping = 1) and beta for novel ones."""
    ov = overlapping.to(pred_open_logits.dtype)
    p, m = pred_open_logits.softmax(dim=-1), mask_pred_open_logits.softmax(dim=-1)
    base = (p ** (1 - alpha) * m ** alpha).log() * ov
    novel = (p ** (1 - beta) * m ** beta).log() * (1 - ov)
    return base + novel


def merge_with_void(pred_logits, open_logits):
    """CategoryODISE.forward, clip_head without bg labels (odise.py:300-323): class distribution from the ensemble,
    void probability from the category head -> log-probabilities [B, Q, K+1]."""
    p_void = F.softmax(pred_logits, dim=-1)[..., -1:]
    cls = F.softmax(open_logits, dim=-1) * (1 - p_void)
    return torch.log(torch.cat([cls, p_void], dim=-1) + 1e-8)


# ---------------------------------------------------------------------------------------------------- text tower
# CLIP text tower (open_clip 2.0.2 `CLIP.encode_text`; glue vendored as ClipAdapter._encode_text, clip.py:138-152 —
# pinned in tests/test_oracle_cpu.py).  The same module with the projection skipped restates SD-v1's
# `cond_stage_model` (ldm FrozenCLIPEmbedder -> HF CLIPTextModel.last_hidden_state; parameter names differ only).

class TextTransformer(nn.Module):
    def __init__(self, vocab=49408, ctx=77, width=768, layers=12, heads=12, out_dim=768):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, width)
        self.positional_embedding = nn.Parameter(torch.zeros(ctx, width))
        self.transformer = Transformer(width, layers, heads)
        self.ln_final = nn.LayerNorm(width)
        self.text_projection = nn.Parameter(torch.zeros(width, out_dim))
        self.logit_scale = nn.Parameter(torch.zeros(()))
        mask = torch.empty(ctx, ctx).fill_(float("-inf")).triu_(1)          # open_clip build_attention_mask
        self.register_buffer("attn_mask", mask, persistent=False)


def encode_text(model, text):
    """ClipAdapter._encode_text (clip.py:138-152): token ids [N, ctx] -> (text_embed [N, out], encodings [N, ctx, width]);
    the embedding is read at the EOT position = argmax of the ids."""
    x = model.token_embedding(text) + model.positional_embedding
    x = model.transformer(x.permute(1, 0, 2), attn_mask=model.attn_mask).permute(1, 0, 2)
    x = model.ln_final(x)
    return x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ model.text_projection, x
