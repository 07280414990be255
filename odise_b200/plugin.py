"""The reference's plugin surface, backed by the B200 engines.

`B200FeatureExtractorBackbone` and `B200MaskFormerHead` keep the forward contracts of
odise.modeling.backbone.FeatureExtractorBackbone (feature_extractor.py:29-256) and
mask2former MaskFormerHead + ODISEMultiScaleMaskedTransformerDecoder (mask_former_head.py:115-132,
odise.py:713-727): NCHW tensors in, the same dict keys out — so a LazyConfig can point `model.backbone` /
`model.sem_seg_head` at them.  NCHW <-> NHWC conversion happens once at this boundary (odise_nchw_to_nhwc_f32).
detectron2 is not required: `output_shape()` returns lightweight ShapeSpec objects with .channels / .stride.
"""
from collections import OrderedDict, namedtuple

import torch
import torch.nn as nn

from . import ops
from .backbone import BackboneEngine
from .head import HeadEngine

ShapeSpec = namedtuple("ShapeSpec", ["channels", "height", "width", "stride"])


class B200FeatureExtractorBackbone(nn.Module):
    """Drop-in for FeatureExtractorBackbone(feature_extractor=LdmImplicitCaptionerExtractor(...), slide_training=True).
    forward(img: [B, 3, H, W] in [0, 1], H, W % 64 == 0) -> {"s2".."s5": [B, 512, H/2^k, W/2^k]}."""

    def __init__(self, state_dict, device, out_features=("s2", "s3", "s4", "s5"), nmma=3, with_vae=True,
                 with_clip=True):
        super().__init__()
        vae = None
        if with_vae:
            from .vae import VAEEngine
            vae = VAEEngine(state_dict, device, nmma=nmma)
        clip = None
        if with_clip:                            # ClipAdapter image tower (clip.py:177-231), weights `clip.visual.*`
            from .clip import ClipVisualEngine
            clip = ClipVisualEngine(state_dict, device, nmma=nmma)
        self.engine = BackboneEngine(state_dict, device, nmma=nmma, vae=vae, clip=clip)
        self._out_features = list(out_features)
        self._out_feature_strides = {f"s{k}": 2 ** k for k in (2, 3, 4, 5)}
        self._out_feature_channels = {f"s{k}": 512 for k in (2, 3, 4, 5)}

    @property
    def size_divisibility(self):
        return 64                                # feature_extractor.py:126-128

    def output_shape(self):
        return {n: ShapeSpec(self._out_feature_channels[n], None, None, self._out_feature_strides[n])
                for n in self._out_features}

    def ignored_state_dict(self, destination=None, prefix=""):
        return OrderedDict() if destination is None else destination     # frozen SD / CLIP weights: nothing to save

    @torch.no_grad()
    def forward(self, img):
        if not img.is_cuda:
            raise RuntimeError("B200FeatureExtractorBackbone: CUDA tensor required (no CPU path)")
        B, _, H, W = img.shape
        if H % 64 or W % 64:
            raise RuntimeError("input must be padded to a multiple of size_divisibility=64")
        feats = self.engine.forward(B, H, W, images_u8=img.contiguous().float())
        return {k: ops.nhwc_to_nchw(t, B, h, w) for k, (t, h, w) in feats.items() if k in self._out_features}


class B200MaskFormerHead(nn.Module):
    """Drop-in for MaskFormerHead(pixel_decoder=MSDeformAttnPixelDecoder, transformer_predictor=
    ODISEMultiScaleMaskedTransformerDecoder(class_embed=PseudoClassEmbed, post_mask_embed=PooledMaskEmbed))."""

    def __init__(self, state_dict, device, num_classes=133, num_queries=100, nmma=3):
        super().__init__()
        self.engine = HeadEngine(state_dict, device, nmma=nmma, num_queries=num_queries)
        self.num_classes = num_classes          # set by OpenPanopticInference through open_state_dict (odise.py:135)
        self._pseudo_classes = num_classes      # PseudoClassEmbed.num_classes is NOT updated by the wrapper

    def _pseudo_logits(self, B, Q, dev):
        # PseudoClassEmbed (odise.py:910-920): ones for every class, zero for background
        lg = torch.ones(B, Q, self._pseudo_classes + 1, device=dev)
        lg[..., -1] = 0
        return lg

    @torch.no_grad()
    def forward(self, features, mask=None):
        B = features["s2"].shape[0]
        feats = {k: (ops.nchw_to_nhwc(v.float()), v.shape[2], v.shape[3]) for k, v in features.items()}
        out = self.engine.forward(feats, B)
        h2, w2 = out["pd"]["mask_hw"]
        Q = self.engine.Q
        dev = features["s2"].device

        def pack(hd):
            return dict(pred_logits=self._pseudo_logits(B, Q, dev), pred_masks=hd["pred_masks"].view(B, Q, h2, w2),
                        mask_embed=hd["mask_embed"].view(B, Q, -1),
                        mask_pooled_features=hd["mask_pooled_features"].view(B, Q, -1),
                        logit_scale=torch.tensor(self.engine.logit_scale, device=dev))

        res = pack(out["heads"][-1])
        res["aux_outputs"] = [pack(h) for h in out["heads"][:-1]]
        return res

    def layers(self, features, mask=None):
        return self.forward(features, mask)


class B200PoolingCLIPHead(nn.Module):
    """Drop-in for PoolingCLIPHead (odise.py:1420-1542), inference only, normalize_logits=True, no bg labels:
    forward(outputs) pops "pred_open_logits" [B, Q, K] and returns {"pred_open_logits": ensemble logits [B, Q, K]}
    from outputs["images"] ([B, 3, H, W] in [0, 1]) and outputs["pred_masks"].  The vocabulary (CLIP text embeddings of
    the prompts, prompt counts per class, overlap with the training vocabulary) is set with set_vocabulary()."""

    def __init__(self, state_dict, device, alpha=0.35, beta=0.65, nmma=3, visual=None):
        super().__init__()
        from .clip import ClipVisualEngine, MaskClipHead
        self.visual = visual if visual is not None else ClipVisualEngine(state_dict, device, nmma=nmma)
        import math
        self.engine = MaskClipHead(self.visual, alpha=alpha, beta=beta,
                                   logit_scale=math.exp(float(state_dict.get("clip.logit_scale", math.log(100.0)))))
        self.alpha, self.beta = alpha, beta

    @property
    def with_bg(self):
        return False

    def set_vocabulary(self, text_embed, group_sizes, overlapping):
        self.engine.set_vocabulary("test", text_embed, group_sizes, overlapping)

    @torch.no_grad()
    def forward(self, outputs, targets=None):
        assert not self.training, "PoolingCLIPHead only supports inference"
        assert targets is None and "test" in self.engine._vocab
        open_logits = outputs.pop("pred_open_logits")
        img, masks = outputs["images"], outputs["pred_masks"]
        if not img.is_cuda:
            raise RuntimeError("B200PoolingCLIPHead: CUDA tensors required (no CPU path)")
        B, Q, K = open_logits.shape
        # the device kernel takes the category logits with a void column; it only enters the merged output, which the
        # plugin surface does not return (the caller merges, odise.py:300-323)
        cat = torch.cat([open_logits.float(), torch.zeros(B, Q, 1, device=img.device)], dim=-1)
        r = self.engine.forward("test", img.contiguous().float(), B, img.shape[2], img.shape[3], masks.float(), cat,
                                want_open=True)
        return {"pred_open_logits": r["pred_open_logits"]}


class B200CategoryODISE(nn.Module):
    """Drop-in for the eval branch of CategoryODISE.forward (odise.py:209-246, :282-370): a list of
    {"image": uint8 [3, h, w] (0..255), "height": H_out, "width": W_out} in, a list of
    {"sem_seg": [K, H_out, W_out], "panoptic_seg": (int32 [H_out, W_out], segments_info), "instances": {...}} out.
    Batching follows detectron2's ImageList.from_tensors: images are top-left aligned in a zero-padded batch whose size
    is the per-batch maximum rounded up to size_divisibility (64) for the network, and un-rounded for MaskCLIP.
    `engine` is an ODISEEngine with a vocabulary set (set_vocabulary / set_vocabulary_from_tokens)."""

    def __init__(self, engine, size_divisibility=64, semantic_on=True, panoptic_on=True, instance_on=True,
                 test_topk_per_image=100, tokenizer=None, train_labels=None, prompt="photo", metadata=None):
        super().__init__()
        self.engine = engine
        self.size_divisibility = size_divisibility
        self.semantic_on, self.panoptic_on, self.instance_on = semantic_on, panoptic_on, instance_on
        self.test_topk_per_image = test_topk_per_image
        self.tokenizer, self.train_labels, self.prompt = tokenizer, train_labels, prompt
        self.metadata = metadata
        self.num_classes = None
        self.test_labels = None

    # ---- vocabulary protocol of OpenPanopticInference (odise/modeling/wrapper/pano_wrapper.py:36-68, odise.py:133-166)
    def open_state_dict(self, destination=None, prefix=""):
        d = OrderedDict() if destination is None else destination
        d[prefix + "sem_seg_head.num_classes"] = self.num_classes
        d[prefix + "metadata"] = self.metadata
        d[prefix + "test_topk_per_image"] = self.test_topk_per_image
        d[prefix + "semantic_on"] = self.semantic_on
        d[prefix + "panoptic_on"] = self.panoptic_on
        d[prefix + "instance_on"] = self.instance_on
        d[prefix + "category_head.test_labels"] = self.test_labels
        d[prefix + "clip_head.test_labels"] = self.test_labels
        return d

    def load_open_state_dict(self, state_dict):
        """Same keys as open_state_dict().  New `test_labels` (list of synonym lists) build — or re-activate from the cache,
        like get_and_cache_test_text_embed — the vocabulary on the device; `metadata.thing_dataset_id_to_contiguous_id`
        names the "thing" classes of the panoptic merge (maskformer_model.py:318)."""
        known = set(self.open_state_dict())
        for k, v in state_dict.items():
            if k not in known:
                raise KeyError(f"{k} is not part of the open state dict")
        g = state_dict.get
        self.num_classes = g("sem_seg_head.num_classes", self.num_classes)
        self.metadata = g("metadata", self.metadata)
        self.test_topk_per_image = g("test_topk_per_image", self.test_topk_per_image)
        self.semantic_on = g("semantic_on", self.semantic_on)
        self.panoptic_on = g("panoptic_on", self.panoptic_on)
        self.instance_on = g("instance_on", self.instance_on)
        cat, clip = g("category_head.test_labels", self.test_labels), g("clip_head.test_labels", self.test_labels)
        if "category_head.test_labels" in state_dict and "clip_head.test_labels" in state_dict and cat != clip:
            raise ValueError("category_head and clip_head must share one test vocabulary")
        labels = cat if "category_head.test_labels" in state_dict else clip
        if labels is not None and labels != self.test_labels:
            self._activate(labels)
        self.test_labels = labels
        for k, v in state_dict.items():                       # the reference asserts every key took (odise.py:166)
            assert self.open_state_dict()[k] == v, f"{k} is not loaded correctly"

    def _activate(self, labels):
        from . import vocab
        key = tuple(tuple(s) for s in labels)
        if self.num_classes is not None and self.num_classes != len(labels):
            raise ValueError(f"num_classes = {self.num_classes} but {len(labels)} test labels")
        if self.engine.has_vocabulary(key):
            self.engine.use_vocabulary(key)
            return
        if self.tokenizer is None:
            raise RuntimeError("a new test vocabulary needs a tokenizer (odise_b200.vocab.SimpleTokenizer)")
        things = None
        if self.metadata is not None and hasattr(self.metadata, "thing_dataset_id_to_contiguous_id"):
            things = sorted(self.metadata.thing_dataset_id_to_contiguous_id.values())
        vocab.build_vocabulary(self.engine, self.tokenizer, key, labels, self.train_labels, things, self.prompt)

    @torch.no_grad()
    def forward(self, batched_inputs):
        assert not self.training, "B200CategoryODISE is inference only"
        eng, dev = self.engine, self.engine.dev
        imgs = [x["image"] for x in batched_inputs]
        sizes = [(int(im.shape[-2]), int(im.shape[-1])) for im in imgs]
        mh, mw = max(s[0] for s in sizes), max(s[1] for s in sizes)
        d = self.size_divisibility
        ph, pw = (mh + d - 1) // d * d, (mw + d - 1) // d * d
        n = len(imgs)
        net = torch.zeros(n, 3, ph, pw, dtype=torch.uint8, device=dev)
        for i, im in enumerate(imgs):
            if im.dtype != torch.uint8:
                raise RuntimeError('"image" must be uint8 CHW in 0..255 (detectron2 DatasetMapper format)')
            net[i, :, :sizes[i][0], :sizes[i][1]] = im.to(dev, non_blocking=True)
        clip_in = net[:, :, :mh, :mw].contiguous() if (mh, mw) != (ph, pw) else net
        out = eng.step(n, ph, pw, images_u8=net, clip_images=clip_in)
        results = []
        for i, x in enumerate(batched_inputs):
            H, W = int(x.get("height", sizes[i][0])), int(x.get("width", sizes[i][1]))
            post = eng.post(out["pred_logits"][i:i + 1], out["pred_masks"][i:i + 1], H, W, semantic=self.semantic_on,
                            panoptic=self.panoptic_on, instance=self.instance_on, topk=self.test_topk_per_image,
                            panoptic_on=self.panoptic_on, padded_size=(ph, pw), image_size=sizes[i])
            r = {}
            if self.semantic_on:
                r["sem_seg"] = post["sem_seg"][0]
            if self.panoptic_on:
                r["panoptic_seg"] = (post["panoptic_seg"][0], eng.post.segments_info(post["seg_info"], post["n_segments"])[0])
            if self.instance_on:
                ins = post["instances"]
                keep = ins["valid"][0].bool()
                qi = ins["query_index"][0][keep].long()
                r["instances"] = dict(pred_masks=ins["query_masks"][0][qi], scores=ins["scores"][0][keep],
                                      pred_classes=ins["pred_classes"][0][keep].long(), image_size=(H, W))
            results.append(r)
        return results
