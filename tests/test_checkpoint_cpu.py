"""Real-weight loading (odise_b200/checkpoint.py) against files laid out like the reference's three
downloads (odise/checkpoint/odise_checkpointer.py:54-140): "state_dict" payload for SD, TorchScript / plain CLIP, "model"
payload for ODISE with numpy entries and optimizer state alongside."""
import numpy as np
import pytest
import torch

from odise_b200 import checkpoint as ck
from odise_b200 import spec


def _small_expected():
    return (spec.clip_visual_params(width=64, layers=1, patch=14, image=28, out_dim=32)
            + spec.clip_text_params(width=64, layers=1, vocab=50, ctx=9, out_dim=32)
            + spec.sd_text_params(width=64, layers=1, vocab=50, ctx=9)
            + spec.category_head_params()
            + [("model.diffusion_model.time_embed.0.weight", (8, 4), "w"), ("first_stage_model.quant_conv.weight", (8, 8, 1, 1), "w"),
               ("sem_seg_head.predictor.query_feat.weight", (10, 16), "emb")])


def _write(tmp_path):
    exp = _small_expected()
    sd = spec.synth_state_dict(exp, 0)
    ldm = {k: v for k, v in sd.items() if k.startswith(ck.LDM_PREFIXES)}
    ldm["model_ema.decay"] = torch.tensor(0.999)
    ldm["cond_stage_model.transformer.text_model.embeddings.position_ids"] = torch.arange(9).view(1, -1)
    torch.save({"state_dict": ldm, "global_step": 7}, tmp_path / "sd.ckpt")
    clip = {k[len("clip."):]: v for k, v in sd.items() if k.startswith("clip.")}
    clip["input_resolution"] = torch.tensor(28)
    torch.save(clip, tmp_path / "clip.pt")
    od = {k: v for k, v in sd.items() if k.startswith(("sem_seg_head.", "category_head."))}
    od["category_head.text_proj.bias"] = od["category_head.text_proj.bias"].numpy()          # detectron2-style ndarray
    od["criterion.empty_weight"] = torch.ones(3)
    torch.save({"model": od, "optimizer": {"state": {}}, "iteration": 5}, tmp_path / "odise.pth")
    return exp, sd


def test_roundtrip_three_files(tmp_path):
    exp, sd = _write(tmp_path)
    got = ck.assemble(ck.read_ldm_checkpoint(tmp_path / "sd.ckpt"), ck.read_clip_checkpoint(tmp_path / "clip.pt"),
                      ck.read_odise_checkpoint(tmp_path / "odise.pth"))
    rep = ck.verify(got, exp)
    assert rep["missing"] == [] and rep["mismatched"] == [] and rep["unexpected"] == []
    for n, _, _ in exp:
        assert torch.equal(got[n].float(), sd[n]), n
    assert isinstance(got["category_head.text_proj.bias"], torch.Tensor)


def test_incompatible_files_fail_loudly(tmp_path):
    exp, sd = _write(tmp_path)
    got = ck.assemble(ck.read_ldm_checkpoint(tmp_path / "sd.ckpt"), None, ck.read_odise_checkpoint(tmp_path / "odise.pth"))
    with pytest.raises(ck.CheckpointError, match="missing"):
        ck.verify(got, exp)
    rep = ck.verify(got, exp, strict=False)
    assert "clip.visual.conv1.weight" in rep["missing"]
    bad = dict(ck.assemble(got, ck.read_clip_checkpoint(tmp_path / "clip.pt")))
    bad["clip.visual.proj"] = torch.zeros(3, 3)
    bad["totally.unknown"] = torch.zeros(1)
    with pytest.raises(ck.CheckpointError, match="shape mismatches"):
        ck.verify(bad, exp)
    assert ck.verify(bad, exp, strict=False)["unexpected"] == ["totally.unknown"]
    with pytest.raises(ck.CheckpointError):
        ck.read_odise_checkpoint(tmp_path / "sd.ckpt")             # wrong file in the wrong slot
    with pytest.raises(ck.CheckpointError):
        ck.read_clip_checkpoint(tmp_path / "odise.pth")
    with pytest.raises(ck.CheckpointError, match="conflicting"):
        ck.assemble({"a": torch.zeros(2)}, {"a": torch.ones(2)})


class Payload:                                                      # a pickled object, like lightning callbacks
    pass


def test_untrusted_pickles_are_refused(tmp_path):
    torch.save({"state_dict": {"model.diffusion_model.x": torch.zeros(1)}, "callbacks": Payload()}, tmp_path / "pl.ckpt")
    with pytest.raises(ck.CheckpointError, match="trusted=True"):
        ck.read_ldm_checkpoint(tmp_path / "pl.ckpt")
    assert "model.diffusion_model.x" in ck.read_ldm_checkpoint(tmp_path / "pl.ckpt", trusted=True)


def test_full_inventory_is_consistent():
    exp = ck.expected_params()
    names = [n for n, _, _ in exp]
    assert len(names) == len(set(names))
    n_par = sum(int(np.prod(s)) for _, s, _ in exp)
    assert 1.4e9 < n_par < 1.6e9          # SD-v1 UNet 860M + VAE 84M + CLIP-L/14-336 428M + SD text 123M + ODISE head
