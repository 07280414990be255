"""Inference meta-arch of the B200 engine: the eval branch of CategoryODISE.forward
(odise/modeling/meta_arch/odise.py:209-246, :282-331) restricted to the north-star hot path:

    images -> [sliding 512^2 crops] -> UNet feature pass -> projections -> pixel decoder -> masked-attention decoder
           -> mask_embed x CLIP-text scoring -> (pred_logits [B, Q, K+1], pred_masks [B, Q, H/4, W/4])

One process per GPU; images shard over ranks with a single NCCL all-gather of the final logits (reference analogue:
d2 evaluator gather).  A step at fixed (batch, H, W) is captured once into a CUDA graph and replayed.
"""
import math

import torch

from . import lib, spec
from .backbone import BackboneEngine
from .head import HeadEngine


def full_param_list(with_vae=False, with_clip=False):
    return (spec.unet_params() + spec.backbone_params() + spec.head_params() + (spec.vae_params() if with_vae else [])
            + (spec.clip_visual_params() if with_clip else []))


class ODISEEngine:
    def __init__(self, sd, device, nmma=3, num_queries=100, with_vae=False, with_clip=False, with_clip_head=None,
                 alpha=0.3, beta=0.7, uncond=None):
        """with_clip: CLIP ViT-L/14-336 image tower (implicit captioner input, §8f-2); with_clip_head (default = with_clip):
        MaskCLIP + PoolingCLIPHead ensemble (odise_with_label.py: alpha 0.3, beta 0.7) on the same frozen tower."""
        self.dev = torch.device(device)
        self.nmma = nmma
        vae = None
        if with_vae:                      # real KL-VAE taps instead of synthetic ones
            from .vae import VAEEngine
            vae = VAEEngine(sd, device, nmma=nmma)
        clip = None
        if with_clip:                     # real CLIP ViT-L/14-336 image embedding of every crop
            from .clip import ClipVisualEngine
            clip = ClipVisualEngine(sd, device, nmma=nmma)
        self.with_vae, self.with_clip = with_vae, with_clip
        self.clip_head = None
        if with_clip_head if with_clip_head is not None else with_clip:
            from .clip import MaskClipHead
            assert clip is not None, "the MaskCLIP head shares the CLIP image tower: with_clip=True required"
            self.clip_head = MaskClipHead(clip, alpha=alpha, beta=beta,
                                          logit_scale=math.exp(float(sd.get("clip.logit_scale", math.log(100.0)))))
        self.backbone = BackboneEngine(sd, device, nmma=nmma, vae=vae, clip=clip, uncond=uncond)
        self.text = None                  # ClipTextEngine, built on demand (set_vocabulary_from_tokens)
        self._sd_text = {k: v for k, v in sd.items() if k.startswith(spec.CLIP_TEXT_PREFIX) and not k.startswith(spec.CLIP_PREFIX)}
        self._null_embed = sd.get("category_head.null_embed")
        self.head = HeadEngine(sd, device, nmma=nmma, num_queries=num_queries)
        self.Q = num_queries
        self._graphs = {}
        self.vocab_key = None
        self.launches_per_step = None

    def set_vocabulary(self, key, text_bank, null_bank, group_sizes, thing_ids=None, clip_text_bank=None,
                       overlapping=None):
        """OpenPanopticInference / CategoryEmbed.test_labels analogue (pano_wrapper.py:58-68, odise.py:1281-1307):
        the vocabulary is a [K', 768] CLIP text bank + per-class prompt counts (+ which classes are "things" for
        the panoptic merge, metadata.thing_dataset_id_to_contiguous_id in the reference).  With a MaskCLIP head:
        clip_text_bank = the raw CLIP text embeddings of the same prompts (default: text_bank), overlapping[k] = class k
        shares a name with the training vocabulary (odise.py:1483-1493; default: every other class)."""
        self.head.set_vocabulary(key, text_bank, null_bank, group_sizes)
        if self.clip_head is not None:
            K = len(group_sizes)
            ov = overlapping if overlapping is not None else [(k % 2) == 0 for k in range(K)]
            self.clip_head.set_vocabulary(key, clip_text_bank if clip_text_bank is not None else text_bank, group_sizes, ov)
        from .postprocess import PostProcessor
        K = len(group_sizes)
        if not hasattr(self, "_posts"):
            self._posts = {}
        self._posts[key] = PostProcessor(self.dev, K, thing_ids if thing_ids is not None else range(0, K, 2), nmma=self.nmma)
        self.use_vocabulary(key)

    def has_vocabulary(self, key):
        return key in getattr(self, "_posts", {})

    def use_vocabulary(self, key):
        """Switch to a vocabulary that was set before (the reference caches text banks per label tuple, odise.py:1281-1288)."""
        if not self.has_vocabulary(key):
            raise lib.OdiseError(f"vocabulary {key!r} has not been set")
        self.vocab_key = key
        self.post = self._posts[key]

    @classmethod
    def from_checkpoints(cls, ldm_path, odise_path, clip_path, device, trusted=False, **kw):
        """The reference's three downloads (sd-v1-3.ckpt, OpenAI CLIP ViT-L-14-336, odise_*.pth) -> a ready engine:
        weights verified against the spec inventory, `uncond_inputs` computed by the SD text encoder (ldm.py:116)."""
        from . import checkpoint
        from .clip import uncond_inputs
        sd, _ = checkpoint.load_reference_checkpoints(ldm_path, odise_path, clip_path, trusted=trusted)
        nmma = kw.get("nmma", 3)
        un = uncond_inputs(sd, device, nmma=nmma).cpu()
        return cls(sd, device, with_vae=True, with_clip=True, uncond=un, **kw)

    @torch.no_grad()
    def set_vocabulary_from_tokens(self, key, token_ids, group_sizes, null_token_ids=None, thing_ids=None, overlapping=None):
        """CategoryEmbed.get_and_cache_test_text_embed (odise.py:1281-1288): tokenised prompts [K', 77] (all synonyms of
        all classes, class-major) -> CLIP text bank on the device -> set_vocabulary.  The null embedding is the
        checkpoint's `category_head.null_embed` parameter, or the text embedding of `null_token_ids` ("" at init)."""
        from .clip import ClipTextEngine, build_text_bank
        if self.text is None:
            if not self._sd_text:
                raise lib.OdiseError("no CLIP text tower weights (`clip.token_embedding.weight`, ...) in the state dict")
            self.text = ClipTextEngine(self._sd_text, self.dev, nmma=self.nmma)
        bank = build_text_bank(self.text, token_ids)
        if self._null_embed is not None:
            null = self._null_embed.view(1, -1)
        elif null_token_ids is not None:
            null = self.text.encode(null_token_ids.view(1, -1))[0]
        else:
            raise lib.OdiseError("no category_head.null_embed in the state dict and no null_token_ids given")
        self.set_vocabulary(key, bank, null, group_sizes, thing_ids=thing_ids, clip_text_bank=bank, overlapping=overlapping)
        return bank

    @torch.no_grad()
    def postprocess(self, out, H, W, semantic=True, panoptic=True, instance=False, padded_size=None, image_size=None):
        """semantic / panoptic / instance inference on the device (odise.py:326-370).  (H, W) = requested output size;
        padded_size / image_size as in sem_seg_postprocess when the output is not the padded input itself."""
        return self.post(out["pred_logits"], out["pred_masks"], H, W, semantic=semantic, panoptic=panoptic,
                         instance=instance, padded_size=padded_size, image_size=image_size)

    # ------------------------------------------------------------------------------------------- device step
    @torch.no_grad()
    def step(self, n_images, H, W, vae_taps=None, images_u8=None, clip_images=None):
        """One pass of the hot path for n_images resident images (eager). Returns device tensors.
        clip_images: what MaskCLIP looks at when it differs from the network input — the reference feeds it the batch
        padded only to its own maximum size, not to size_divisibility (odise.py:240-246)."""
        if images_u8 is None and (self.with_vae or self.with_clip):
            images_u8 = self._image_buffer(n_images, H, W)
        feats = self.backbone.forward(n_images, H, W, vae_taps, images_u8)
        out = self.head.forward(feats, n_images, vocab_key=self.vocab_key)
        h2, w2 = out["pd"]["mask_hw"]
        last = out["heads"][-1]
        res = dict(pred_masks=last["pred_masks"].view(n_images, self.Q, h2, w2),
                   mask_embed=last["mask_embed"].view(n_images, self.Q, -1),
                   mask_pooled_features=last["mask_pooled_features"].view(n_images, self.Q, -1))
        if "pred_logits" in out:
            res["pred_logits"] = out["pred_logits"]
            if self.clip_head is not None:        # odise.py:292-323: MaskCLIP ensemble replaces the class scores
                ci = clip_images if clip_images is not None else images_u8
                ch = self.clip_head.forward(self.vocab_key, ci, n_images, ci.shape[2], ci.shape[3], res["pred_masks"],
                                            out["pred_logits"])
                res["pred_logits_category"] = out["pred_logits"]
                res["pred_logits"] = ch["pred_logits"]
                res["clip_mask_embed"] = ch["mask_embed"]
        res["aux"] = out["heads"][:-1]
        return res

    def _image_buffer(self, n_images, H, W):
        shp = (n_images, 3, H, W)
        if not hasattr(self, "_img_dev") or tuple(self._img_dev.shape) != shp:
            g = torch.Generator().manual_seed(99)
            self._img_dev = torch.randint(0, 256, shp, generator=g, dtype=torch.uint8).to(self.dev)
            self._host_logits = None
        return self._img_dev

    def capture(self, n_images, H, W):
        """Warm up eagerly, then capture the step into a CUDA graph (static shapes, static buffers)."""
        key = (n_images, H, W, self.vocab_key)
        if key in self._graphs:
            return self._graphs[key]
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                self.step(n_images, H, W)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        before = lib.launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self.step(n_images, H, W)
        self.launches_per_step = lib.launch_count() - before
        self._graphs[key] = (g, out)
        return self._graphs[key]

    # ------------------------------------------------------------------------------------------- user API
    @torch.no_grad()
    def infer(self, images_u8_pinned, use_graph=True):
        """End-to-end call on a host batch: uint8 [B, 3, H, W] in pinned memory -> host dict
        (pred_logits [B, Q, K+1], pred_masks [B, Q, H/4, W/4]).  H2D and D2H inside."""
        B, _, H, W = images_u8_pinned.shape
        self._image_buffer(B, H, W).copy_(images_u8_pinned, non_blocking=True)
        if not hasattr(self, "_host_logits"):
            self._host_logits = None
        if use_graph:
            g, out = self.capture(B, H, W)
            g.replay()
        else:
            out = self.step(B, H, W, images_u8=self._img_dev)
        if self._host_logits is None:
            self._host_logits = torch.empty(out["pred_logits"].shape, dtype=torch.float32).pin_memory()
            self._host_masks = torch.empty(out["pred_masks"].shape, dtype=torch.float32).pin_memory()
        self._host_logits.copy_(out["pred_logits"], non_blocking=True)
        self._host_masks.copy_(out["pred_masks"], non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return dict(pred_logits=self._host_logits, pred_masks=self._host_masks)


def gather_logits(local_logits):
    """Image-sharded inference: one all-gather of the final class logits over the ranks (NCCL on NVLink)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_logits
    out = torch.empty((dist.get_world_size() * local_logits.shape[0],) + tuple(local_logits.shape[1:]),
                      dtype=local_logits.dtype, device=local_logits.device)
    dist.all_gather_into_tensor(out, local_logits.contiguous())
    return out


def synthetic_vocabulary(n_classes, n_prompts, seed=11):
    """Seeded stand-in for the CLIP text bank of a vocabulary with n_classes classes and n_prompts prompt strings
    (ADE-150: 150 / 403, COCO-133: 133 / 254, ADE-847: 847 / 1342) -> (bank, null, group sizes)."""
    g = torch.Generator().manual_seed(seed)
    bank = torch.randn(n_prompts, 768, generator=g)
    null = torch.randn(1, 768, generator=torch.Generator().manual_seed(13))
    base, extra = divmod(n_prompts, n_classes)
    sizes = [base + (1 if i < extra else 0) for i in range(n_classes)]
    return bank, null, sizes
