"""The reference plugin surface on top of the engines: NCHW in, same dict keys out."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def _label_head(cuda):
    """B200MaskFormerHead built with the keyword arguments of configs/common/models/mask_generator_with_label.py:29-66."""
    from odise_b200.plugin import (B200MaskFormerHead, B200MSDeformAttnPixelDecoder, B200ODISEMultiScaleMaskedTransformerDecoder,
                                   B200PooledMaskEmbed, B200PseudoClassEmbed)
    return B200MaskFormerHead(
        ignore_value=255, num_classes=133,
        pixel_decoder=B200MSDeformAttnPixelDecoder(conv_dim=256, mask_dim=256, norm="GN", transformer_dropout=0.0,
                                                   transformer_nheads=8, transformer_dim_feedforward=1024,
                                                   transformer_enc_layers=6, transformer_in_features=["s3", "s4", "s5"],
                                                   common_stride=4),
        loss_weight=1.0, transformer_in_feature="multi_scale_pixel_decoder",
        transformer_predictor=B200ODISEMultiScaleMaskedTransformerDecoder(
            class_embed=B200PseudoClassEmbed(num_classes=133), hidden_dim=256,
            post_mask_embed=B200PooledMaskEmbed(hidden_dim=256, mask_dim=256, projection_dim=256),
            in_channels=256, mask_classification=True, num_classes=133, num_queries=100, nheads=8, dim_feedforward=2048,
            dec_layers=9, pre_norm=False, enforce_input_project=False, mask_dim=256),
        device=cuda)


def test_head_plugin_contract(cuda):
    """Constructor keywords of the label config, reference-keyed state dict (`pixel_decoder.*`, `predictor.*`), and the
    sub-module surface of MaskFormerHead.layers (mask_former_head.py:118-120)."""
    from odise_b200 import spec
    from oracle import m2f
    sd = spec.synth_state_dict(spec.head_params(), seed=1)
    head = _label_head(cuda)
    with pytest.raises(RuntimeError):                                  # the engine is built from the weights
        head({"s2": torch.zeros(1, 512, 32, 32, device=cuda)})
    ref_keyed = {k[len("sem_seg_head."):]: v for k, v in sd.items() if k.startswith("sem_seg_head.")}
    with pytest.raises(RuntimeError):                                  # strict: a missing key is an error
        head.load_state_dict({k: v for k, v in ref_keyed.items() if k != "predictor.query_feat.weight"})
    head.load_state_dict(ref_keyed)
    assert set(head.state_dict()) == set(ref_keyed)
    g = torch.Generator().manual_seed(5)
    feats = {f"s{i}": torch.randn(1, 512, 128 // 2 ** i, 128 // 2 ** i, generator=g) for i in (2, 3, 4, 5)}
    dfe = {k: v.to(cuda) for k, v in feats.items()}
    out = head(dfe)
    assert set(out) == {"pred_logits", "pred_masks", "aux_outputs", "mask_embed", "mask_pooled_features", "logit_scale"}
    assert out["pred_logits"].shape == (1, 100, 134) and out["pred_masks"].shape == (1, 100, 32, 32)
    assert len(out["aux_outputs"]) == 9 and out["logit_scale"].dim() == 0
    assert out["pred_logits"][..., :-1].eq(1).all() and out["pred_logits"][..., -1].eq(0).all()
    with torch.no_grad():
        mf, enc0, ms = m2f.pixel_decoder(sd, feats, "sem_seg_head.pixel_decoder.")
        _, masks = m2f.transformer_decoder(sd, ms, mf, "sem_seg_head.predictor.")
    assert _rel(out["aux_outputs"][0]["pred_masks"].cpu(), masks[0]) < 1e-3
    # .pixel_decoder.forward_features (msdeformattn.py:314-358): (mask_features, out[0], multi_scale_features)
    mask_features, tr_enc, multi = head.pixel_decoder.forward_features(dfe)
    assert mask_features.shape == (1, 256, 32, 32) and [tuple(t.shape[2:]) for t in multi] == [(4, 4), (8, 8), (16, 16)]
    assert _rel(mask_features.cpu(), mf) < 1e-3 and _rel(tr_enc.cpu(), enc0) < 1e-3
    for a, b in zip(multi, ms):
        assert _rel(a.cpu(), b) < 1e-3
    # .predictor(x, mask_features, mask) on tensors the caller owns (clones: forces the NCHW -> token-major path)
    out2 = head.predictor([t.clone() for t in multi], mask_features.clone(), None)
    assert _rel(out2["aux_outputs"][0]["pred_masks"].cpu(), masks[0]) < 1e-3
    assert torch.equal(out2["pred_masks"], out["pred_masks"])          # same arithmetic either way
    # from_state_dict keeps the checkpoint-keyed convenience path
    from odise_b200.plugin import B200MaskFormerHead
    h2 = B200MaskFormerHead.from_state_dict(sd, cuda, num_classes=133)
    assert torch.equal(h2(dfe)["pred_masks"], out["pred_masks"])


def test_backbone_plugin_contract(cuda):
    """Constructor keywords of configs/common/models/odise_with_label.py:16-29 + reference-keyed load_state_dict."""
    from odise_b200 import spec
    from odise_b200.plugin import B200FeatureExtractorBackbone, B200LdmImplicitCaptionerExtractor
    frozen = spec.synth_state_dict(spec.unet_params() + spec.vae_params() + spec.clip_visual_params(), seed=0)
    learn = spec.synth_state_dict(spec.backbone_params(), seed=2)
    bb = B200FeatureExtractorBackbone(
        feature_extractor=B200LdmImplicitCaptionerExtractor(
            encoder_block_indices=(5, 7), unet_block_indices=(2, 5, 8, 11), decoder_block_indices=(2, 5), steps=(0,),
            learnable_time_embed=True, num_timesteps=1, clip_model_name="ViT-L-14-336",
            frozen_state_dict=frozen, synthetic_uncond=True, device=cuda),
        out_features=["s2", "s3", "s4", "s5"], use_checkpoint=True, slide_training=True)
    assert bb.size_divisibility == 64 and bb.feature_extractor.feature_dims == (512, 512, 2560, 1920, 960, 640, 512, 512)
    assert bb.feature_extractor.grouped_indices == [[i] for i in range(8)] and bb.feature_extractor.num_groups == 8
    shp = bb.output_shape()
    assert [shp[k].stride for k in ("s2", "s3", "s4", "s5")] == [4, 8, 16, 32] and shp["s2"].channels == 512
    ref_keyed = {k[len("backbone."):]: v for k, v in learn.items()}
    assert any(k.startswith("feature_projections.0.0.conv1.") for k in ref_keyed) and "feature_extractor.alpha_cond" in ref_keyed
    with pytest.raises(RuntimeError):
        bb(torch.rand(1, 3, 512, 512, device=cuda))                   # no weights yet
    bb.load_state_dict(ref_keyed)
    assert set(bb.state_dict()) == set(ref_keyed) and len(bb.ignored_state_dict()) == 0
    img = torch.rand(1, 3, 512, 512, generator=torch.Generator().manual_seed(2)).to(cuda)
    out = bb(img)
    assert {k: tuple(v.shape) for k, v in out.items()} == {"s2": (1, 512, 128, 128), "s3": (1, 512, 64, 64),
                                                           "s4": (1, 512, 32, 32), "s5": (1, 512, 16, 16)}
    assert all(torch.isfinite(v).all() for v in out.values())
    with pytest.raises(RuntimeError):
        bb(img.cpu())
    with pytest.raises(NotImplementedError):                           # anything but the released configuration is refused
        B200LdmImplicitCaptionerExtractor(unet_block_indices=(2, 5, 8))


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
