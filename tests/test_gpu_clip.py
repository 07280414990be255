"""GPU parity of the CLIP ViT-L/14-336 image-tower engine (odise_b200/clip.py) vs oracle/clip.py
(glue pinned against the reference's ClipAdapter._encode_image in tests/test_oracle_cpu.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def test_clip_preprocess_and_patchify(cuda):
    from odise_b200 import ops
    from oracle import clip as oclip
    g = torch.Generator().manual_seed(3)
    img = torch.rand(2, 3, 640, 768, generator=g)
    boxes = torch.tensor([[0, 0, 0], [1, 128, 256], [0, 64, 200]], dtype=torch.int32)
    got = ops.clip_preprocess(img.to(cuda), boxes.to(cuda), 3, 640, 768, 512, 512, 336).view(3, 336, 336, 3).cpu()
    for i, (im, y, x) in enumerate(boxes.tolist()):
        want = oclip.preprocess(img[im:im + 1, :, y:y + 512, x:x + 512], 336)[0].permute(1, 2, 0)
        assert (got[i] - want).abs().max() < 2e-5
    u8 = torch.randint(0, 256, (1, 3, 512, 512), generator=g, dtype=torch.uint8)
    got = ops.clip_preprocess(u8.to(cuda), torch.tensor([[0, 0, 0]], dtype=torch.int32).to(cuda), 1, 512, 512, 512, 512, 336)
    want = oclip.preprocess(u8.float() / 255.0, 336)[0].permute(1, 2, 0).reshape(-1, 3)
    assert (got.cpu() - want).abs().max() < 2e-5
    x = torch.randn(2, 28, 28, 3, generator=g)
    p = ops.patchify_split(x.to(cuda).reshape(-1, 3), 2, 28, 14)
    ref = torch.nn.functional.unfold(x.permute(0, 3, 1, 2), 14, stride=14).transpose(1, 2).reshape(8, 588)
    assert _rel(p.float()[:, :588].cpu(), ref) < 1e-4 and p.float()[:, 588:].abs().max() == 0


def test_clip_image_embed(cuda):
    from odise_b200 import spec
    from odise_b200.clip import ClipVisualEngine
    from oracle import clip as oclip
    sd = spec.synth_state_dict(spec.clip_visual_params(), seed=5)
    with torch.device("meta"):
        v = oclip.VisionTransformer()
    v.load_state_dict({k[len(spec.CLIP_PREFIX):]: t for k, t in sd.items()}, assign=True)
    v.eval()
    g = torch.Generator().manual_seed(9)
    img = torch.rand(1, 3, 512, 1024, generator=g)
    boxes = torch.tensor([[0, 0, 0], [0, 0, 512]], dtype=torch.int32)
    with torch.no_grad():
        want = torch.cat([oclip.embed_image(v, img[:, :, :, :512]), oclip.embed_image(v, img[:, :, :, 512:])])
    eng = ClipVisualEngine(sd, cuda, nmma=3)
    got = eng.embed(img.to(cuda), boxes.to(cuda), 2, 512, 1024, 512, 512)
    torch.cuda.synchronize()
    assert got.shape == (2, 768)
    assert _rel(got.cpu(), want) < 1e-3, _rel(got.cpu(), want)
