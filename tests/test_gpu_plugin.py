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
