"""GPU parity of the backbone glue (odise_b200/backbone.py): implicit-captioner conditioning (ldm.py:705-714),
q_sample at t=0 (gaussian_diffusion.py:275-292), BottleneckBlock projections summed per stride
(feature_extractor.py:157-179) and the crop paste / average of slide_forward (feature_extractor.py:205-248)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def bb(cuda):
    from odise_b200 import spec
    from odise_b200.backbone import BackboneEngine
    sd = spec.synth_state_dict(spec.unet_params() + spec.backbone_params(), seed=0)
    return sd, BackboneEngine(sd, cuda, nmma=3, synthetic_uncond=True)


def test_conditioning_and_q_sample(cuda, bb):
    from oracle import ldm
    sd, eng = bb
    B = 3
    g = torch.Generator().manual_seed(2)
    emb = torch.randn(B, 768, generator=g)
    ctx, cemb = eng.conditioning(emb.to(cuda), B)
    e = "backbone.feature_extractor."
    uncond = torch.randn(1, 77, 768, generator=torch.Generator().manual_seed(17))
    proj = torch.nn.functional.linear(emb.double(), sd[e + "clip_project.linear.weight"].double(), sd[e + "clip_project.linear.bias"].double())
    prefix = proj.unsqueeze(1) + sd[e + "clip_project.positional_embedding"].double()
    want = uncond.double() + torch.tanh(sd[e + "alpha_cond"].double()) * prefix
    assert _rel(ctx.view(B, 77, 768).cpu(), want) < 1e-4
    tp = torch.nn.functional.linear(emb.double(), sd[e + "time_embed_project.linear.weight"].double(), sd[e + "time_embed_project.linear.bias"].double())
    want = torch.tanh(sd[e + "alpha_cond_time_embed"].double()) * (tp.unsqueeze(1) + sd[e + "time_embed_project.positional_embedding"].double())
    assert _rel(cemb.view(B, 1, 1280).cpu(), want) < 1e-4
    lat = torch.randn(B, 4, 64, 64, generator=g)
    got = eng.q_sample(lat.permute(0, 2, 3, 1).reshape(-1, 4).contiguous().to(cuda), B, 64, 64)
    want = ldm.q_sample_t0(lat, ldm.shared_noise())
    assert _rel(got.view(B, 64, 64, 4).permute(0, 3, 1, 2).cpu(), want) < 1e-6


def test_projections_vs_oracle(cuda, bb):
    from oracle import m2f
    from odise_b200.backbone import TAP_ORDER, FEATURE_DIMS
    sd, eng = bb
    B, H = 1, 256
    g = torch.Generator().manual_seed(4)
    native = [4, 8, 64, 32, 16, 8, 8, 4]        # un-clamped strides of the taps (ldm.py:284-346)
    feats = [torch.randn(B, FEATURE_DIMS[i], H // native[i], H // native[i], generator=g) for i in range(8)]
    with torch.no_grad():
        want = m2f.forward_features(sd, feats, (H, H))
    taps = {n: (f.permute(0, 2, 3, 1).reshape(-1, f.shape[1]).contiguous().to(cuda), f.shape[2], f.shape[3])
            for n, f in zip(TAP_ORDER, feats)}
    got = eng.project(taps, B, (H, H))
    torch.cuda.synchronize()
    for k, w in want.items():
        t, h, ww = got[k]
        assert _rel(t.view(B, h, ww, 512).permute(0, 3, 1, 2).cpu(), w) < 1e-3, k


def test_slide_paste_average(cuda, bb):
    """paste + count-average against the reference rule restated in torch (overlapping 3x3 crops of a 640^2 image
    with 256^2 crops: same code path as 1280^2 / 512^2)."""
    sd, eng = bb
    n_img, Himg, crop = 2, 640, 256
    boxes, short = eng.crop_grid(Himg, Himg, crop)
    assert short == 256 and len(boxes) == 9
    g = torch.Generator().manual_seed(6)
    B = n_img * len(boxes)
    per = {k: torch.randn(B, crop // s, crop // s, 512, generator=g) for k, s in (("s2", 4), ("s3", 8), ("s4", 16), ("s5", 32))}
    eng_extract = eng.extract
    try:
        eng.extract = lambda B_, hw, *a, **kw: {k: (v.reshape(-1, 512).contiguous().to(cuda), v.shape[1], v.shape[2]) for k, v in per.items()}
        eng_crop_grid = eng.crop_grid
        eng.crop_grid = lambda h, w, c=512: eng_crop_grid(h, w, crop)
        out = eng.forward(n_img, Himg, Himg)
    finally:
        eng.extract = eng_extract
        eng.crop_grid = eng_crop_grid
    torch.cuda.synchronize()
    for k, s in (("s2", 4), ("s3", 8), ("s4", 16), ("s5", 32)):
        Hd = Himg // s
        want = torch.zeros(n_img, Hd, Hd, 512)
        cnt = torch.zeros(Hd, Hd)
        for img in range(n_img):
            for ci, (y1, x1) in enumerate(boxes):
                fh = crop // s
                want[img, y1 // s:y1 // s + fh, x1 // s:x1 // s + fh] += per[k][img * 9 + ci]
                if img == 0:
                    cnt[y1 // s:y1 // s + fh, x1 // s:x1 // s + fh] += 1
        want = want / cnt[None, :, :, None]
        t, h, w = out[k]
        assert (h, w) == (Hd, Hd)
        assert _rel(t.view(n_img, Hd, Hd, 512).cpu(), want) < 1e-6, k
