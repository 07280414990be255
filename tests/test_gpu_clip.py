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


def test_maskclip_preprocess_and_bits(cuda):
    import torch.nn.functional as F
    from odise_b200 import ops
    from oracle import clip as oclip
    g = torch.Generator().manual_seed(11)
    img = torch.rand(2, 3, 320, 448, generator=g)
    got = ops.maskclip_preprocess(img.to(cuda), 2, 320, 448, 336).view(2, 336, 336, 3).permute(0, 3, 1, 2).cpu()
    want = oclip.preprocess(F.interpolate(img, size=(336, 336), mode="bilinear", align_corners=False), 336)
    assert (got - want).abs().max() < 2e-5
    u8 = torch.randint(0, 256, (1, 3, 1024, 1024), generator=g, dtype=torch.uint8)
    got = ops.maskclip_preprocess(u8.to(cuda), 1, 1024, 1024, 336).view(1, 336, 336, 3).permute(0, 3, 1, 2).cpu()
    want = oclip.preprocess(F.interpolate(u8.float() / 255.0, size=(336, 336), mode="bilinear", align_corners=False), 336)
    assert (got - want).abs().max() < 2e-5
    # attention bits of the mask tokens vs the reference's bool mask (True = blocked)
    with torch.device("meta"):
        v = oclip.VisionTransformer()
    B, Q, T, TS = 2, 9, 577, 592
    masks = torch.randn(B, Q, 64, 80, generator=g) * 2 - 1.5
    masks[1, 3] = -4.0                                                     # touches no patch: only the class key stays
    am = oclip.mask_attention_mask(v, F.interpolate(masks, size=(336, 336), mode="bilinear", align_corners=False))
    am = am.view(B, 16, Q + T, Q + T)[:, 0, :Q, Q:]                          # [B, Q, 577] blocked flags of mask rows
    bits, row_any = ops.maskclip_bits(masks.to(cuda), B, Q, 64, 80, 336, 14, TS, T)
    bits, row_any = bits.cpu(), row_any.cpu()
    assert row_any[:, T:T + Q].eq(1).all() and row_any[:, :T].eq(0).all() and row_any[:, T + Q:].eq(0).all()
    keys = torch.arange(T)
    on = ((bits[:, T:T + Q][..., keys // 32] >> (keys % 32)) & 1).bool()
    assert torch.equal(on, ~am)
    assert on[1, 3].sum() == 1 and on[..., 0].all()


def test_open_vocab_merge(cuda):
    from odise_b200 import ops
    from oracle import clip as oclip
    g = torch.Generator().manual_seed(12)
    rows, K = 37, 150
    cat = torch.randn(rows, K + 1, generator=g) * 6
    clip = torch.randn(rows, K + 1, generator=g) * 8                       # last column = padding (row stride K+1)
    ov = (torch.rand(K, generator=g) < 0.5)
    merged, op = ops.open_vocab_merge(cat.to(cuda), clip.to(cuda), K + 1, ov.to(torch.uint8).to(cuda), 0.3, 0.7, rows, K,
                                      want_open=True)
    ens = oclip.pooling_clip_ensemble(cat[:, :K].double(), clip[:, :K].double(), ov.long(), 0.3, 0.7)
    want = oclip.merge_with_void(cat.double(), ens)
    assert (op.cpu().double() - ens).abs().max() < 1e-4
    assert (merged.cpu().double() - want).abs().max() < 1e-4


def test_maskclip_head(cuda):
    """MaskCLIP.get_mask_embed + pred_logits + PoolingCLIPHead ensemble + void merge vs the oracle (full ViT-L/14-336)."""
    from odise_b200 import spec
    from odise_b200.clip import ClipVisualEngine, MaskClipHead
    from oracle import clip as oclip
    sd = spec.synth_state_dict(spec.clip_visual_params(), seed=5)
    with torch.device("meta"):
        v = oclip.VisionTransformer()
    v.load_state_dict({k[len(spec.CLIP_PREFIX):]: t for k, t in sd.items()}, assign=True)
    v.eval()
    g = torch.Generator().manual_seed(13)
    N, Q, K = 1, 12, 6
    img = torch.rand(N, 3, 512, 512, generator=g)
    yy, xx = torch.meshgrid(torch.arange(128).float(), torch.arange(128).float(), indexing="ij")
    masks = torch.stack([(20 + 4 * q - ((yy - 10 * q) ** 2 + (xx - 64) ** 2).sqrt()) * 0.5 for q in range(Q)])[None]
    masks = masks + torch.randn(N, Q, 128, 128, generator=g) * 0.2
    sizes = [2, 1, 3, 1, 1, 2]
    text = torch.randn(sum(sizes), 768, generator=g)
    cat = torch.randn(N, Q, K + 1, generator=g) * 3
    ov = torch.tensor([1, 0, 1, 1, 0, 0])
    with torch.no_grad():
        me = oclip.get_mask_embed(v, img, masks)
        lg = oclip.maskclip_pred_logits(me, text, sizes, 100.0)
        want = oclip.merge_with_void(cat, oclip.pooling_clip_ensemble(cat[..., :-1], lg, ov, 0.3, 0.7))
    eng = ClipVisualEngine(sd, cuda, nmma=3)
    head = MaskClipHead(eng, alpha=0.3, beta=0.7, logit_scale=100.0)
    head.set_vocabulary("v", text, sizes, ov)
    out = head.forward("v", img.to(cuda), N, 512, 512, masks.to(cuda), cat.to(cuda))
    torch.cuda.synchronize()
    assert _rel(out["mask_embed"].view(N, Q, -1).cpu(), me) < 1e-3, _rel(out["mask_embed"].view(N, Q, -1).cpu(), me)
    got_lg = out["mask_pred_open_logits"].view(N, Q, K + 1)[..., :K].cpu()
    assert (got_lg - lg).abs().max() < 2e-2          # logits are 100 * cos-sim: 2e-4 of their scale
    assert (out["pred_logits"].cpu() - want).abs().max() < 2e-2


def test_clip_text_tower_and_uncond(cuda):
    """CLIP text tower (open_clip names) vs oracle.encode_text (== ClipAdapter._encode_text, pinned), and the SD-v1
    cond_stage_model (HF names) producing uncond_inputs for the empty prompt (ldm.py:116)."""
    from odise_b200 import spec
    from odise_b200.clip import ClipTextEngine, build_text_bank, uncond_inputs, EMPTY_PROMPT_IDS
    from oracle import clip as oclip
    sd = spec.synth_state_dict(spec.clip_text_params(), seed=8)
    with torch.device("meta"):
        m = oclip.TextTransformer()
    m.load_state_dict({k[len(spec.CLIP_TEXT_PREFIX):]: v for k, v in sd.items()}, assign=True)
    m.attn_mask = torch.empty(77, 77).fill_(float("-inf")).triu_(1)
    m.eval()
    g = torch.Generator().manual_seed(4)
    ids = torch.zeros(5, 77, dtype=torch.int64)
    for i, n in enumerate((2, 5, 9, 30, 77)):                               # prompt lengths incl. BOS/EOT; 77 = full
        ids[i, :n] = torch.randint(1000, 40000, (n,), generator=g)
        ids[i, 0], ids[i, n - 1] = 49406, 49407
    with torch.no_grad():
        emb, enc = oclip.encode_text(m, ids)
    eng = ClipTextEngine(sd, cuda, nmma=3)
    got_emb, got_enc = eng.encode(ids)
    torch.cuda.synchronize()
    assert _rel(got_emb.cpu(), emb) < 1e-3 and _rel(got_enc.cpu(), enc) < 1e-3
    assert torch.equal(build_text_bank(eng, ids, batch=2).cpu(), got_emb.cpu())
    # SD text encoder: HF parameter names, no projection
    hf = spec.synth_state_dict(spec.sd_text_params(), seed=9)
    conv = spec.hf_text_to_openai(hf, dst_prefix="")
    with torch.device("meta"):
        t = oclip.TextTransformer()
    t.load_state_dict(conv, assign=True, strict=False)
    t.attn_mask = m.attn_mask
    t.text_projection = torch.nn.Parameter(torch.zeros(768, 768))
    with torch.no_grad():
        _, want = oclip.encode_text(t.eval(), torch.tensor([EMPTY_PROMPT_IDS]))
    got = uncond_inputs(hf, cuda)
    assert got.shape == (1, 77, 768) and _rel(got.cpu(), want) < 1e-3
