"""B200 engine for the CLIP ViT-L/14-336 image tower behind ClipAdapter.embed_image (SURVEY.md §8f-2;
odise/modeling/meta_arch/clip.py:177-231, called from LdmImplicitCaptionerExtractor.forward, ldm.py:705).

crop [B, 3, 512, 512] in [0,1] -> bicubic 336 + CLIP normalisation (odise_clip_preprocess) -> 14x14 patch embedding as a
GEMM -> 24 pre-LN transformer blocks (QKV / out / MLP on the tcgen05 GEMM with fused bias, QuickGELU and residual
epilogues; attention on the tcgen05 flash kernel, d = 64) -> ln_post + projection of the class token -> [B, 768].
Tokens per image are padded 577 -> 584 rows so every TMA box start stays 16-byte aligned; pad keys are masked.
"""
import torch

from . import lib, ops, spec
from .lib import Planes
from .ops import ACT_QUICKGELU


class ClipVisualEngine:
    def __init__(self, sd, device, nmma=3, prefix=spec.CLIP_PREFIX, width=1024, layers=24, heads=16, patch=14, image=336):
        self.dev = torch.device(device)
        self.nmma, self.lo = nmma, nmma == 3
        self.width, self.layers, self.heads, self.patch, self.image = width, layers, heads, patch, image
        self.G = image // patch
        self.T = self.G * self.G + 1             # 577
        self.TS = (self.T + 7) // 8 * 8          # 584 rows per image
        f = lambda t: t.to(self.dev, torch.float32).contiguous()
        g = lambda n: sd[prefix + n]

        def pl(w):
            w = f(w)
            if w.shape[1] % 8:
                w = torch.nn.functional.pad(w, (0, 8 - w.shape[1] % 8))
            return lib.split(w, lo=self.lo)

        self.W, self.F = {}, {}
        self.W["conv1"] = pl(g("conv1.weight").reshape(width, -1))
        pos = g("positional_embedding").float()
        self.F["pos_patches"] = f(pos[1:])                                   # [576, width]
        self.F["cls_row"] = f((g("class_embedding").float() + pos[0]).view(1, width))
        self.F["ln_pre.g"], self.F["ln_pre.b"] = f(g("ln_pre.weight")), f(g("ln_pre.bias"))
        for i in range(layers):
            q = f"transformer.resblocks.{i}."
            n = f"l{i}."
            w, b = g(q + "attn.in_proj_weight"), g(q + "attn.in_proj_bias")
            self.W[n + "qk"], self.F[n + "qk.b"] = pl(w[:2 * width]), f(b[:2 * width])
            self.W[n + "v"], self.F[n + "v.b"] = pl(w[2 * width:]), f(b[2 * width:])
            self.W[n + "o"], self.F[n + "o.b"] = pl(g(q + "attn.out_proj.weight")), f(g(q + "attn.out_proj.bias"))
            self.W[n + "fc"], self.F[n + "fc.b"] = pl(g(q + "mlp.c_fc.weight")), f(g(q + "mlp.c_fc.bias"))
            self.W[n + "pr"], self.F[n + "pr.b"] = pl(g(q + "mlp.c_proj.weight")), f(g(q + "mlp.c_proj.bias"))
            for ln in ("ln_1", "ln_2"):
                self.F[n + ln + ".g"], self.F[n + ln + ".b"] = f(g(q + ln + ".weight")), f(g(q + ln + ".bias"))
        self.F["ln_post.g"], self.F["ln_post.b"] = f(g("ln_post.weight")), f(g("ln_post.bias"))
        self.W["proj"] = pl(g("proj").t())                                   # x @ proj == x @ (proj^T)^T

    def _gemm(self, a, name, **kw):
        return lib.gemm(a, self.W[name], nmma=self.nmma, bias=self.F.get(name + ".b"), **kw)

    @torch.no_grad()
    def embed(self, img, boxes_dev, n_crops, H, W, ch, cw):
        """img: device uint8 / float32 [N, 3, H, W]; boxes [n_crops, 3] int32 -> image_embed fp32 [n_crops, 768]."""
        dev, B, Wd, T, TS = self.dev, n_crops, self.width, self.T, self.TS
        x = ops.clip_preprocess(img, boxes_dev, B, H, W, ch, cw, self.image)
        patches = ops.patchify_split(x, B, self.image, self.patch, lo=self.lo)
        tok = torch.zeros(B * TS, Wd, dtype=torch.float32, device=dev)
        # patch embedding + positional embedding written straight into rows 1..576 of every image's token block
        lib.gemm(patches, self.W["conv1"], M=T - 1, N=Wd, K=patches.cols, nmma=self.nmma, batch=B,
                 a_bs=(T - 1) * patches.ld, residual=self.F["pos_patches"], ld_res=Wd, res_bs=0,
                 out=tok[1:], ld_out=Wd, out_bs=TS * Wd)
        # class rows: class_embedding + positional_embedding[0]
        ops.copy2d(self.F["cls_row"].expand(B, Wd), tok.view(B, TS * Wd)[:, :Wd])
        h, _ = ops.layer_norm(tok, self.F["ln_pre.g"], self.F["ln_pre.b"], want_f32=True, want_planes=False, lo=self.lo)
        M = B * TS
        d = Wd // self.heads
        for i in range(self.layers):
            n = f"l{i}."
            _, y = ops.layer_norm(h, self.F[n + "ln_1.g"], self.F[n + "ln_1.b"], lo=self.lo)
            qk = Planes.empty(M, 2 * Wd, dev, lo=self.lo)
            self._gemm(y, n + "qk", out_planes=qk)
            vt = Planes.empty(Wd, M, dev, lo=self.lo)
            lib.gemm(self.W[n + "v"], y, nmma=self.nmma, bias_m=self.F[n + "v.b"], out_planes=vt)
            _, o = ops.attention_tc(qk.col_slice(0, Wd), qk.col_slice(Wd, Wd), vt, B, self.heads, d, TS, T, d ** -0.5,
                                    self.nmma, tk_stride=TS)
            h2 = ops.empty(M, Wd, dev)
            self._gemm(o, n + "o", residual=h, out=h2)
            _, y2 = ops.layer_norm(h2, self.F[n + "ln_2.g"], self.F[n + "ln_2.b"], lo=self.lo)
            u = Planes.empty(M, 4 * Wd, dev, lo=self.lo)
            self._gemm(y2, n + "fc", act=ACT_QUICKGELU, out_planes=u)
            h = ops.empty(M, Wd, dev)
            self._gemm(u, n + "pr", residual=h2, out=h)
        cls = h.view(B, TS * Wd)[:, :Wd]                                       # token 0 of every image (strided rows)
        _, c = ops.layer_norm(cls, self.F["ln_post.g"], self.F["ln_post.b"], lo=self.lo)
        out = ops.empty(B, self.W["proj"].rows, dev)
        lib.gemm(c, self.W["proj"], nmma=self.nmma, out=out)
        return out
