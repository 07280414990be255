"""TEST INFRASTRUCTURE (oracle side): the reference's backbone composed from the oracle pieces exactly the way
FeatureExtractorBackbone composes them —
  slide_forward (feature_extractor.py:181-250): crop grid, per-crop single_forward, paste-add + count, divide
  single_forward (feature_extractor.py:139-155): T.Resize((512, 512), BICUBIC) of the crop, feature extractor, forward_features
  LdmImplicitCaptionerExtractor.forward (ldm.py:697-718) -> LdmExtractor.forward (ldm.py:543-621)
Only tests/ use this (never odise_b200/)."""
import torch
import torch.nn.functional as F

from odise_b200 import spec
from oracle import clip as oclip, ldm, m2f


def load_modules(sd):
    def load(cls, prefix):
        with torch.device("meta"):
            m = cls()
        m.load_state_dict({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}, assign=True)
        return m.eval()
    return dict(unet=load(ldm.UNetModel, spec.UNET_PREFIX), vae=load(ldm.AutoencoderKL, spec.VAE_PREFIX),
                vis=load(oclip.VisionTransformer, spec.CLIP_PREFIX))


@torch.no_grad()
def single_forward(sd, mods, crop01, uncond, in_size=512):
    """crop01 [1, 3, h, w] in [0, 1] -> {s2..s5} at (h, w) / stride."""
    h, w = crop01.shape[-2:]
    img = crop01
    if (h, w) != (in_size, in_size):      # image_preprocess = T.Resize((512, 512), BICUBIC); float tensors are not clamped
        img = F.interpolate(crop01, size=(in_size, in_size), mode="bicubic", align_corners=False)
    e = "backbone.feature_extractor."
    lin = F.linear
    emb = oclip.embed_image(mods["vis"], img)                                                    # ldm.py:705
    ctx = uncond + torch.tanh(sd[e + "alpha_cond"]) * (
        lin(emb, sd[e + "clip_project.linear.weight"], sd[e + "clip_project.linear.bias"]).unsqueeze(1)
        + sd[e + "clip_project.positional_embedding"])
    cemb = torch.tanh(sd[e + "alpha_cond_time_embed"]) * (
        lin(emb, sd[e + "time_embed_project.linear.weight"], sd[e + "time_embed_project.linear.bias"]).unsqueeze(1)
        + sd[e + "time_embed_project.positional_embedding"])
    lat, ef = ldm.encoder_features(mods["vae"], (img - 0.5) / 0.5)
    uf = ldm.unet_features(mods["unet"], ldm.q_sample_t0(lat, ldm.shared_noise(lat.shape[-2:])), ctx, cemb[:, 0])
    df = ldm.decoder_features(mods["vae"], lat)
    return m2f.forward_features(sd, [*ef, *uf, *df], (h, w))


@torch.no_grad()
def slide_forward(sd, mods, img01, uncond, crop=512):
    """img01 [1, 3, H, W] -> {s2..s5: [1, 512, H/s, W/s]} (feature_extractor.py:181-250, slide_training=True)."""
    _, _, H, W = img01.shape
    short = min(crop, min(H, W))
    hg = max(H - short + short - 1, 0) // short + 1
    wg = max(W - short + short - 1, 0) // short + 1
    out, cnt = {}, {}
    for hi in range(hg):
        for wi in range(wg):
            y2, x2 = min(hi * short + short, H), min(wi * short + short, W)
            y1, x1 = max(y2 - short, 0), max(x2 - short, 0)
            f = single_forward(sd, mods, img01[:, :, y1:y2, x1:x2], uncond)
            for k, v in f.items():
                s = short // v.shape[-1]
                if k not in out:
                    out[k] = torch.zeros(1, v.shape[1], H // s, W // s)
                    cnt[k] = torch.zeros(1, 1, H // s, W // s)
                out[k][:, :, y1 // s:y2 // s, x1 // s:x2 // s] += v
                cnt[k][:, :, y1 // s:y2 // s, x1 // s:x2 // s] += 1
    return {k: out[k] / cnt[k] for k in out}
