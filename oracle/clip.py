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

    def forward(self, x):                      # [L, N, D]
        y = self.ln_1(x)
        x = x + self.attn(y, y, y, need_weights=False)[0]
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(width, heads) for _ in range(layers)])

    def forward(self, x):
        for r in self.resblocks:
            x = r(x)
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
