"""B200 engine for ODISE's diffusion backbone glue: the sliding 512x512 crops of
FeatureExtractorBackbone.slide_forward (odise/modeling/backbone/feature_extractor.py:181-250), the implicit
captioner front of LdmImplicitCaptionerExtractor.forward (odise/modeling/meta_arch/ldm.py:697-718), q_sample at
t = 0 (odise/modeling/diffusion/gaussian_diffusion.py:275-292) with the reference's shared noise (ldm.py:271-277),
the UNet feature pass (unet.py) and the eight BottleneckBlock projections summed per stride
(feature_extractor.py:157-179).

All crops of all images of a step run as ONE batch through the UNet (B = images x crops): the reference loops over
crops sequentially (feature_extractor.py:205-227).

The KL-VAE encoder / truncated decoder and the CLIP image tower are §8(f) rows; until their engines land, their
outputs enter through a `TapProvider` (synthetic, seeded) — see DESIGN.md.
"""
import math

import torch

from . import lib, ops, spec
from .ops import ACT_RELU
from .unet import UNetEngine, CTX_T

FEATURE_DIMS = spec.FEATURE_DIMS
FEATURE_STRIDES = spec.FEATURE_STRIDES
TAP_ORDER = ("enc5", "enc7", "unet2", "unet5", "unet8", "unet11", "dec2", "dec5")


def t0_coefficients():
    """sqrt(alpha_bar_0), sqrt(1 - alpha_bar_0) of the "ldm_linear" schedule with 1000 steps, rounded to fp32 like
    _extract_into_tensor(...).float() does (gaussian_diffusion.py:125-135, :288-291)."""
    import numpy as np
    betas = np.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=np.float64) ** 2
    ac = np.cumprod(1.0 - betas, axis=0)
    return float(np.float32(np.sqrt(ac)[0])), float(np.float32(np.sqrt(1.0 - ac)[0]))


class SyntheticTaps:
    """Stand-in for the KL-VAE (encoder taps, latent, decoder taps) and the CLIP image embedding: deterministic
    seeded tensors of the reference's shapes (ldm.py:284-346).  NOT part of the measured hot path."""

    def __init__(self, device, seed=21):
        self.dev = torch.device(device)
        self.seed = seed
        self._cache = {}

    def __call__(self, B, crop_hw=(512, 512)):
        key = (B, crop_hw)
        if key not in self._cache:
            g = torch.Generator().manual_seed(self.seed)
            H, W = crop_hw
            mk = lambda h, w, c: torch.randn(B * h * w, c, generator=g).to(self.dev)
            self._cache[key] = dict(
                latent=(mk(H // 8, W // 8, 4), H // 8, W // 8),
                enc5=(mk(H // 4, W // 4, 512), H // 4, W // 4), enc7=(mk(H // 8, W // 8, 512), H // 8, W // 8),
                dec2=(mk(H // 8, W // 8, 512), H // 8, W // 8), dec5=(mk(H // 4, W // 4, 512), H // 4, W // 4),
                clip_embed=torch.randn(B, 768, generator=g).to(self.dev))
        return self._cache[key]


class BackboneEngine:
    def __init__(self, sd, device, nmma=3, prefix="backbone.", unet_prefix=spec.UNET_PREFIX, uncond=None, vae=None, clip=None):
        """sd: state dict with `backbone.feature_projections.*`, `backbone.feature_extractor.*` and the UNet.
        uncond: the frozen text-encoder output for "" ([1, 77, 768]; ldm.py:116) — an input of the path.
        vae: optional VAEEngine; without it the VAE taps / latent are synthetic.
        clip: optional ClipVisualEngine (§8f-2); without it the CLIP image embedding is a seeded synthetic tensor."""
        self.dev = torch.device(device)
        self.nmma, self.lo = nmma, nmma == 3
        self.vae = vae
        self.clip = clip
        self._boxes = {}
        self.unet = UNetEngine(sd, device, nmma=nmma, prefix=unet_prefix)
        self.W, self.F = {}, {}
        f = lambda t: t.to(self.dev, torch.float32).contiguous()
        pl = lambda w: lib.split(f(w.reshape(w.shape[0], -1)), lo=self.lo)
        for i, cin in enumerate(FEATURE_DIMS):
            q = f"{prefix}feature_projections.{i}.0."
            n = f"p{i}."
            self.W[n + "c1"] = pl(sd[q + "conv1.weight"])
            self.W[n + "c2"] = lib.split(f(sd[q + "conv2.weight"].permute(0, 2, 3, 1).reshape(128, -1)), lo=self.lo)
            self.W[n + "c3"] = pl(sd[q + "conv3.weight"])
            for c in ("conv1", "conv2", "conv3"):
                self.F[n + c + ".g"], self.F[n + c + ".b"] = f(sd[q + c + ".norm.weight"]), f(sd[q + c + ".norm.bias"])
            if cin != 512:
                self.W[n + "sc"] = pl(sd[q + "shortcut.weight"])
                self.F[n + "sc.g"], self.F[n + "sc.b"] = f(sd[q + "shortcut.norm.weight"]), f(sd[q + "shortcut.norm.bias"])
        e = prefix + "feature_extractor."
        self.W["clip_project"] = pl(sd[e + "clip_project.linear.weight"])
        self.F["clip_project.b"] = f(sd[e + "clip_project.linear.bias"])
        self.W["time_project"] = pl(sd[e + "time_embed_project.linear.weight"])
        self.F["time_project.b"] = f(sd[e + "time_embed_project.linear.bias"])
        if uncond is None:
            uncond = torch.randn(1, CTX_T, 768, generator=torch.Generator().manual_seed(17))
        # weight-only terms of cond = uncond + tanh(alpha) * (proj + pos)   (ldm.py:707-709)
        ta = torch.tanh(sd[e + "alpha_cond"].float())
        self.F["cond.ta"] = f(ta.view(CTX_T, 768))
        self.F["cond.a0"] = f((uncond.float() + ta * sd[e + "clip_project.positional_embedding"].float()).view(CTX_T, 768))
        tt = torch.tanh(sd[e + "alpha_cond_time_embed"].float()).view(1, 1280)
        self.F["temb.ta"] = f(tt)
        self.F["temb.a0"] = f(tt * sd[e + "time_embed_project.positional_embedding"].float().view(1, 1280))
        c0, c1 = t0_coefficients()
        self.c0 = c0
        noise = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(42))      # ldm.py:273-276
        self.F["noise_c1"] = f((c1 * noise).permute(0, 2, 3, 1).reshape(64 * 64, 4))
        self.taps_provider = SyntheticTaps(device)

    # ------------------------------------------------------------------------------------------- pieces
    def conditioning(self, clip_embed, B):
        """-> (context [B*77, 768], cond_emb [B, 1280])"""
        e_p = ops.split(clip_embed, lo=self.lo)
        proj = ops.empty(B, 768, self.dev)
        lib.gemm(e_p, self.W["clip_project"], nmma=self.nmma, bias=self.F["clip_project.b"], out=proj)
        ctx = ops.bcast_fma(self.F["cond.a0"], self.F["cond.ta"], proj, B, CTX_T, 768)
        tp = ops.empty(B, 1280, self.dev)
        lib.gemm(e_p, self.W["time_project"], nmma=self.nmma, bias=self.F["time_project.b"], out=tp)
        cemb = ops.bcast_fma(self.F["temb.a0"], self.F["temb.ta"], tp, B, 1, 1280)
        return ctx, cemb

    def q_sample(self, latent, B, h, w):
        """x_t = sqrt(abar_0) * z + sqrt(1 - abar_0) * eps with the shared 64x64 noise (64x64 latents only)."""
        if (h, w) != (64, 64):
            raise lib.OdiseError("q_sample: only the 64x64 latent of a 512x512 crop is supported")
        x = ops.empty(B * h * w, 4, self.dev)
        ops.copy2d(latent, x, scale=self.c0)
        y, _ = ops.add_split(x, self.F["noise_c1"], b_rows=h * w, want_f32=True, want_planes=False)
        return y

    def project(self, taps, B, crop_hw):
        """forward_features (feature_extractor.py:157-179): -> {"s2".."s5": (NHWC fp32 [B*h*w, 512], h, w)}"""
        H, W = crop_hw
        out = {}
        for idx, name in enumerate(TAP_ORDER):
            x, h, w = taps[name]
            s = FEATURE_STRIDES[idx]
            th, tw = H // s, W // s
            if (h, w) != (th, tw):
                x = ops.resize_nhwc(x, B, h, w, th, tw, bilinear=False)       # F.interpolate default = nearest
            M = B * th * tw
            n = f"p{idx}."
            x_p = ops.split(x, lo=self.lo)
            t1 = ops.empty(M, 128, self.dev)
            lib.gemm(x_p, self.W[n + "c1"], nmma=self.nmma, out=t1)
            _, a1 = ops.group_norm(t1, B, th * tw, self.F[n + "conv1.g"], self.F[n + "conv1.b"], 1e-5, ACT_RELU, lo=self.lo)
            t2 = ops.empty(M, 128, self.dev)
            lib.gemm(a1, self.W[n + "c2"], M=M, N=128, nmma=self.nmma, conv=(128, th, tw), out=t2)
            _, a2 = ops.group_norm(t2, B, th * tw, self.F[n + "conv2.g"], self.F[n + "conv2.b"], 1e-5, ACT_RELU, lo=self.lo)
            t3 = ops.empty(M, 512, self.dev)
            lib.gemm(a2, self.W[n + "c3"], nmma=self.nmma, out=t3)
            if FEATURE_DIMS[idx] != 512:
                sc = ops.empty(M, 512, self.dev)
                lib.gemm(x_p, self.W[n + "sc"], nmma=self.nmma, out=sc)
                sc, _ = ops.group_norm(sc, B, th * tw, self.F[n + "sc.g"], self.F[n + "sc.b"], 1e-5, want_f32=True,
                                       want_planes=False)
            else:
                sc = x
            key = f"s{int(math.log2(s))}"
            first = key not in out
            if first:
                out[key] = (ops.empty(M, 512, self.dev), th, tw)
            ops.group_norm_res(t3, B, th * tw, self.F[n + "conv3.g"], self.F[n + "conv3.b"], 1e-5, sc, ACT_RELU,
                               out[key][0], accumulate=not first)
        return out

    @torch.no_grad()
    def extract(self, B, crop_hw=(512, 512), vae_taps=None, crops=None, clip_embed=None):
        """single_forward for a batch of B crops.  crops: normalised NHWC fp32 [B*h*w, 3] -> the VAE engine produces
        latent + taps; otherwise they are given / synthetic.  clip_embed: [B, 768] from the CLIP image tower
        (ldm.py:705); synthetic when no ClipVisualEngine is attached."""
        t = vae_taps if vae_taps is not None else self.taps_provider(B, crop_hw)
        if clip_embed is not None:
            t = dict(t, clip_embed=clip_embed)
        if crops is not None and self.vae is not None:
            enc = self.vae.encode(crops, B, crop_hw[0], crop_hw[1])
            lat, lh, lw = enc["latent"]
            dec = self.vae.decode_taps(lat, B, lh, lw)
            t = dict(latent=enc["latent"], enc5=enc["enc5"], enc7=enc["enc7"], dec2=dec["dec2"], dec5=dec["dec5"],
                     clip_embed=t["clip_embed"])
        ctx, cemb = self.conditioning(t["clip_embed"], B)
        lat, lh, lw = t["latent"]
        x = self.q_sample(lat, B, lh, lw)
        u = self.unet.forward(x, B, lh, lw, ctx, cemb)
        taps = dict(enc5=t["enc5"], enc7=t["enc7"], dec2=t["dec2"], dec5=t["dec5"],
                    unet2=u[0], unet5=u[1], unet8=u[2], unet11=u[3])
        return self.project(taps, B, crop_hw)

    # ------------------------------------------------------------------------------------------- sliding window
    @staticmethod
    def crop_grid(h_img, w_img, crop=512):
        """slide_forward's crop boxes (feature_extractor.py:197-218): (y1, x1) list and the crop side."""
        short = min(crop, min(h_img, w_img))
        hg = max(h_img - short + short - 1, 0) // short + 1
        wg = max(w_img - short + short - 1, 0) // short + 1
        boxes = []
        for hi in range(hg):
            for wi in range(wg):
                y2 = min(hi * short + short, h_img)
                x2 = min(wi * short + short, w_img)
                boxes.append((max(y2 - short, 0), max(x2 - short, 0)))
        return boxes, short

    @torch.no_grad()
    def forward(self, n_images, h_img, w_img, vae_taps=None, images_u8=None):
        """slide_forward over n_images images of h_img x w_img: all crops in one batch, paste + average.
        images_u8: device image batch [n_images, 3, H, W], uint8 (0..255) or float32 in [0, 1] (used when a VAE
        engine is attached).
        Returns {"s2".."s5": (NHWC fp32 [n_images * H/s * W/s, 512], H/s, W/s)}."""
        boxes, short = self.crop_grid(h_img, w_img)
        nc = len(boxes)
        B = n_images * nc                                # crop batch, image-major: b = img * nc + crop
        crops = clip_embed = None
        if images_u8 is not None and (self.vae is not None or self.clip is not None):
            key = (n_images, h_img, w_img)
            if key not in self._boxes:
                self._boxes[key] = torch.tensor([[i, y, x] for i in range(n_images) for (y, x) in boxes],
                                                dtype=torch.int32).to(self.dev)
            if self.vae is not None:
                crops = ops.image_crops(images_u8, self._boxes[key], B, h_img, w_img, short, short)
            if self.clip is not None:
                clip_embed = self.clip.embed(images_u8, self._boxes[key], B, h_img, w_img, short, short)
        feats = self.extract(B, (short, short), vae_taps, crops, clip_embed)
        if nc == 1 and short == h_img == w_img:
            return feats
        out = {}
        for k, (f, fh, fw) in feats.items():
            s = short // fh
            Hd, Wd = h_img // s, w_img // s
            dst = torch.zeros(n_images * Hd * Wd, 512, dtype=torch.float32, device=self.dev)
            cnt = torch.zeros(Hd, Wd)
            for ci, (y1, x1) in enumerate(boxes):
                cnt[y1 // s:y1 // s + fh, x1 // s:x1 // s + fw] += 1
            for img in range(n_images):
                for ci, (y1, x1) in enumerate(boxes):
                    b = img * nc + ci
                    src = f[b * fh * fw:(b + 1) * fh * fw].view(fh, fw * 512)
                    d0 = (img * Hd + y1 // s) * Wd + x1 // s
                    dv = dst[d0:d0 + (fh - 1) * Wd + fw].as_strided((fh, fw * 512), (Wd * 512, 1))
                    ops.copy2d(src, dv, accumulate=True)
            if float(cnt.max()) > 1:
                inv = (1.0 / cnt).reshape(-1).repeat(n_images).to(self.dev)
                ops.rowscale(dst, inv)
            out[k] = (dst, Hd, Wd)
        return out
