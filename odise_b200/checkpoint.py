"""Real-weight loading: the three files a reference deployment reads, flattened into the ONE
state dict (reference key names, odise_b200/spec.py) the engines consume.

  * sd-v1-*.ckpt           LdmCheckpointer (odise/checkpoint/odise_checkpointer.py:130-140): payload under "state_dict";
                           `model.diffusion_model.*` (UNet), `first_stage_model.*` (KL-VAE), `cond_stage_model.*` (text
                           encoder that produces `uncond_inputs`, ldm.py:116);
  * OpenAI CLIP ViT-L-14-336 (open_clip "openai": a TorchScript archive or a plain state dict): `visual.*` + text tower,
                           re-keyed under `clip.` (the reference keeps two frozen copies, backbone.feature_extractor.clip
                           and clip_head.clip, of the same weights; one is enough here);
  * odise_*.pth            ODISECheckpointer (odise_checkpointer.py:54-127): payload under "model"; only the trainable
                           parts are stored — the frozen SD / CLIP modules return empty state_dict()s
                           (helper.py:35-46, clip.py:107-122) and are reported by the reference as ignored missing keys.

verify() plays the role of the checkpointer's incompatible-keys report: missing / unexpected / shape-mismatched names
against the spec inventory, so a wrong or truncated file fails loudly instead of running on garbage.
"""
import numpy as np
import torch

from . import spec

LDM_PREFIXES = ("model.diffusion_model.", "first_stage_model.", "cond_stage_model.")
# keys of real checkpoints that no engine needs (buffers / EMA shadows / loss state)
IGNORABLE = ("model_ema.", "betas", "alphas_cumprod", "sqrt_", "log_one_minus", "posterior_", "logvar",
             "cond_stage_model.transformer.text_model.embeddings.position_ids", "criterion.", "pixel_mean", "pixel_std",
             "clip.attn_mask", "clip.input_resolution", "clip.context_length", "clip.vocab_size",
             "first_stage_model.loss.")


class CheckpointError(RuntimeError):
    pass


def _tensors(d):
    """_convert_ndarray_to_tensor (detectron2 checkpoints may hold numpy arrays); non-array entries are dropped."""
    out = {}
    for k, v in d.items():
        if isinstance(v, np.ndarray):
            v = torch.from_numpy(v)
        if isinstance(v, torch.Tensor):
            out[k] = v.detach()
    return out


def _load(path, trusted):
    """torch.load restricted to tensors/containers unless the caller vouches for the file (pytorch-lightning
    checkpoints such as sd-v1-*.ckpt pickle callback objects and need trusted=True)."""
    if trusted:
        return torch.load(path, map_location="cpu", weights_only=False)
    # detectron2-style checkpoints store some entries as numpy arrays: allow exactly the array constructors
    try:
        import numpy._core.multiarray as ma            # numpy >= 2
    except ImportError:
        import numpy.core.multiarray as ma             # numpy 1.x
    try:
        allow = [ma._reconstruct, ma.scalar, np.ndarray, np.dtype] + [type(np.dtype(t)) for t in
                 (np.float16, np.float32, np.float64, np.int32, np.int64, np.uint8, np.bool_)]
        with torch.serialization.safe_globals(allow):
            return torch.load(path, map_location="cpu", weights_only=True)
    except Exception as e:  # noqa  (unpickling errors only: import problems above surface as themselves)
        raise CheckpointError(f"{path}: not loadable with weights_only=True ({type(e).__name__}: {e}); "
                              f"pass trusted=True if the file comes from a source you trust") from e


def read_ldm_checkpoint(path, trusted=False):
    ck = _load(path, trusted)
    sd = ck.get("state_dict", ck.get("model", ck)) if isinstance(ck, dict) else ck
    sd = _tensors(sd)
    out = {k: v for k, v in sd.items() if k.startswith(LDM_PREFIXES)}
    if not out:
        raise CheckpointError(f"{path}: no `model.diffusion_model.*` / `first_stage_model.*` keys (not an SD-v1 checkpoint?)")
    return out


def read_clip_checkpoint(path, trusted=False):
    """OpenAI CLIP weights: TorchScript archive (what open_clip downloads for pretrained="openai") or a state dict.
    -> keys under `clip.` (`clip.visual.*`, `clip.transformer.*`, `clip.token_embedding.weight`, ...)."""
    sd = None
    if trusted:       # a TorchScript archive executes code on load: only for files the caller vouches for
        try:
            sd = torch.jit.load(path, map_location="cpu").state_dict()
        except Exception:  # noqa
            sd = None
    if sd is None:
        try:
            ck = _load(path, trusted)
        except CheckpointError as e:
            raise CheckpointError(f"{e} (OpenAI's CLIP download is a TorchScript archive: it needs trusted=True)") from e
        sd = ck.get("state_dict", ck) if isinstance(ck, dict) else ck.state_dict()
    sd = _tensors(sd)
    if "visual.conv1.weight" not in sd:
        raise CheckpointError(f"{path}: no `visual.conv1.weight` (not a CLIP ViT checkpoint?)")
    return {"clip." + k: v for k, v in sd.items()}


def read_odise_checkpoint(path, trusted=False):
    ck = _load(path, trusted)
    sd = _tensors(ck["model"] if isinstance(ck, dict) and "model" in ck else ck)
    if not any(k.startswith("sem_seg_head.") for k in sd):
        raise CheckpointError(f"{path}: no `sem_seg_head.*` keys (not an ODISE checkpoint?)")
    return sd


def assemble(ldm=None, clip=None, odise=None):
    """Merge the per-file dicts; a name present in two files with different contents is an error."""
    out = {}
    for part in (ldm, clip, odise):
        for k, v in (part or {}).items():
            if k in out and (out[k].shape != v.shape or not torch.equal(out[k], v)):
                raise CheckpointError(f"conflicting definitions of {k}")
            out[k] = v
    return out


def expected_params(with_vae=True, with_clip=True, with_text=True):
    ps = spec.unet_params() + spec.backbone_params() + spec.head_params() + [("category_head.null_embed", (1, 768), "pos")]
    if with_vae:
        ps += spec.vae_params()
    if with_clip:
        ps += spec.clip_visual_params()
    if with_text:
        ps += spec.clip_text_params() + spec.sd_text_params()
    return ps


def verify(sd, expected, strict=True):
    """-> dict(missing, unexpected, mismatched); raises CheckpointError when strict and anything required is absent or
    has the wrong shape (unexpected keys alone never raise: real checkpoints carry EMA shadows, schedules, ...)."""
    want = {n: tuple(s) for n, s, _ in expected}
    missing = sorted(n for n in want if n not in sd)
    mismatched = sorted((n, tuple(sd[n].shape), want[n]) for n in want if n in sd and tuple(sd[n].shape) != want[n])
    unexpected = sorted(k for k in sd if k not in want and not any(k.startswith(p) or p in k for p in IGNORABLE))
    rep = dict(missing=missing, unexpected=unexpected, mismatched=mismatched)
    if strict and (missing or mismatched):
        def head(xs):
            return ", ".join(str(x) for x in xs[:5]) + (f", ... (+{len(xs) - 5})" if len(xs) > 5 else "")
        raise CheckpointError(f"incompatible checkpoint: {len(missing)} missing [{head(missing)}]; "
                              f"{len(mismatched)} shape mismatches [{head(mismatched)}]")
    return rep


def load_reference_checkpoints(ldm_path, odise_path, clip_path=None, trusted=False, strict=True):
    """The reference's three downloads -> (state dict for ODISEEngine / plugin classes, verify() report)."""
    sd = assemble(read_ldm_checkpoint(ldm_path, trusted), read_clip_checkpoint(clip_path, trusted) if clip_path else None,
                  read_odise_checkpoint(odise_path, trusted))
    rep = verify(sd, expected_params(with_clip=clip_path is not None, with_text=clip_path is not None), strict)
    return sd, rep
