"""The reference plugin surface on top of the engines: NCHW in, same dict keys out."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def test_head_plugin_contract(cuda):
    from odise_b200 import spec
    from odise_b200.plugin import B200MaskFormerHead
    from oracle import m2f
    sd = spec.synth_state_dict(spec.head_params(), seed=1)
    head = B200MaskFormerHead(sd, cuda, num_classes=133)
    g = torch.Generator().manual_seed(5)
    feats = {f"s{i}": torch.randn(1, 512, 128 // 2 ** i, 128 // 2 ** i, generator=g) for i in (2, 3, 4, 5)}
    out = head({k: v.to(cuda) for k, v in feats.items()})
    assert set(out) == {"pred_logits", "pred_masks", "aux_outputs", "mask_embed", "mask_pooled_features", "logit_scale"}
    assert out["pred_logits"].shape == (1, 100, 134) and out["pred_masks"].shape == (1, 100, 32, 32)
    assert len(out["aux_outputs"]) == 9 and out["logit_scale"].dim() == 0
    assert out["pred_logits"][..., :-1].eq(1).all() and out["pred_logits"][..., -1].eq(0).all()
    with torch.no_grad():
        mf, _, ms = m2f.pixel_decoder(sd, feats, "sem_seg_head.pixel_decoder.")
        _, masks = m2f.transformer_decoder(sd, ms, mf, "sem_seg_head.predictor.")
    assert _rel(out["aux_outputs"][0]["pred_masks"].cpu(), masks[0]) < 1e-3


def test_backbone_plugin_contract(cuda):
    from odise_b200 import spec
    from odise_b200.plugin import B200FeatureExtractorBackbone
    sd = spec.synth_state_dict(spec.unet_params() + spec.backbone_params() + spec.vae_params() +
                                 spec.clip_visual_params(), seed=0)
    bb = B200FeatureExtractorBackbone(sd, cuda)
    assert bb.size_divisibility == 64
    shp = bb.output_shape()
    assert [shp[k].stride for k in ("s2", "s3", "s4", "s5")] == [4, 8, 16, 32] and shp["s2"].channels == 512
    img = torch.rand(1, 3, 512, 512, generator=torch.Generator().manual_seed(2)).to(cuda)
    out = bb(img)
    assert {k: tuple(v.shape) for k, v in out.items()} == {"s2": (1, 512, 128, 128), "s3": (1, 512, 64, 64),
                                                           "s4": (1, 512, 32, 32), "s5": (1, 512, 16, 16)}
    assert all(torch.isfinite(v).all() for v in out.values())
    with pytest.raises(RuntimeError):
        bb(img.cpu())


def test_pooling_clip_head_plugin(cuda):
    """B200PoolingCLIPHead.forward(outputs) == PoolingCLIPHead.forward over MaskCLIP (odise.py:1469-1542) on a small ViT."""
    from odise_b200 import spec
    from odise_b200.clip import ClipVisualEngine
    from odise_b200.plugin import B200PoolingCLIPHead
    from oracle import clip as oclip
    cfg = dict(width=128, layers=2, patch=14, image=56)
    sd = spec.synth_state_dict(spec.clip_visual_params(out_dim=32, **cfg), 5)
    vis = oclip.VisionTransformer(image_size=56, patch=14, width=128, layers=2, heads=2, out_dim=32).eval()
    vis.load_state_dict({k[len(spec.CLIP_PREFIX):]: v for k, v in sd.items()})
    g = torch.Generator().manual_seed(31)
    img = torch.rand(2, 3, 64, 96, generator=g)
    masks = torch.randn(2, 7, 16, 24, generator=g) * 3
    sizes, ov = [2, 1, 2, 1], [1, 0, 0, 1]
    text = torch.randn(sum(sizes), 32, generator=g)
    open_logits = torch.randn(2, 7, 4, generator=g) * 3
    with torch.no_grad():
        lg = oclip.maskclip_pred_logits(oclip.get_mask_embed(vis, img, masks), text, sizes, 100.0)
        want = oclip.pooling_clip_ensemble(open_logits, lg, torch.tensor(ov), 0.35, 0.65)
    head = B200PoolingCLIPHead(sd, cuda, alpha=0.35, beta=0.65, visual=ClipVisualEngine(sd, cuda, nmma=3, heads=2, **cfg)).eval()
    head.set_vocabulary(text, sizes, ov)
    outputs = {"pred_open_logits": open_logits.to(cuda), "images": img.to(cuda), "pred_masks": masks.to(cuda)}
    got = head(outputs)["pred_open_logits"]
    assert "pred_open_logits" not in outputs                      # popped like the reference does
    assert got.shape == (2, 7, 4) and (got.cpu() - want).abs().max() < 2e-2
    with pytest.raises(RuntimeError):
        head({"pred_open_logits": open_logits, "images": img, "pred_masks": masks})
