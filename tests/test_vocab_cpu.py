"""Vocabulary front end (odise_b200/vocab.py): label files / prompts / overlap rule pinned against the reference's own
functions and data (skipped without /root/reference), the BPE tokenizer checked on a hand-built merge table."""
import os

import pytest
import torch

from odise_b200 import vocab
from oracle import refshim

needs_ref = pytest.mark.skipif(not refshim.available(), reason="/root/reference not present")
LABELS = "/root/reference/odise/data/datasets/openseg_labels"


@needs_ref
def test_label_files_and_prompts_match_reference():
    import importlib
    refshim.install()
    rb = importlib.import_module("odise.data.build")
    for name, n_cls, n_prompts in (("ade20k_150", 150, 403), ("coco_panoptic", 133, 254), ("ade20k_847", 847, 1342)):
        mine = vocab.read_label_file(os.path.join(LABELS, f"{name}_with_prompt_eng.txt"))
        ref = rb.get_openseg_labels(name, prompt_engineered=True)
        assert mine == ref and len(mine) == n_cls and sum(len(s) for s in mine) == n_prompts     # SURVEY.md §8 K' counts
        for p in (None, "a", "photo", "scene"):
            assert vocab.prompt_labels(mine, p) == rb.prompt_labels(ref, p)
    assert vocab.read_label_file(os.path.join(LABELS, "ade20k_150.txt")) == rb.get_openseg_labels("ade20k_150")


def test_overlap_rule():
    test_labels = [["cat", "kitty"], ["unicorn"], ["dog"], ["spaceship", "rocket"]]
    train_labels = [["cat"], ["dog", "puppy"], ["tree"]]
    assert vocab.overlapping_mask(test_labels, train_labels) == [True, False, True, False]      # as pinned in ref_clip.pt
    with pytest.raises(ValueError):
        vocab.prompt_labels(test_labels, "poem")
    assert vocab.prompt_labels([["wall"]], "photo") == [["a photo of a wall."]]


def test_bpe_tokenizer_mechanics():
    """merge table: 'c a' -> 'ca', 'ca t</w>' -> 'cat</w>', 'd o' -> 'do', 'p h' -> 'ph'.  Ranks decide the merge order;
    unmerged characters stay single tokens; the last character carries '</w>'."""
    tk = vocab.SimpleTokenizer(merges=["c a", "ca t</w>", "d o", "p h"])
    e = tk.encoder
    assert len(e) == 256 + 256 + 4 + 2 and tk.sot_id == len(e) - 2 and tk.eot_id == len(e) - 1
    assert tk.encode("cat") == [e["cat</w>"]]
    assert tk.encode("Cat  ") == [e["cat</w>"]]                                  # lower-cased, whitespace cleaned
    assert tk.encode("cats") == [e["ca"], e["t"], e["s</w>"]]                     # 'ca t</w>' needs the word-final t
    assert tk.encode("dog") == [e["do"], e["g</w>"]]
    assert tk.encode("a photo.") == [e["a</w>"], e["ph"], e["o"], e["t"], e["o</w>"], e[".</w>"]]
    assert tk.encode("it's 42") == [e["i"], e["t</w>"], e["'"], e["s</w>"], e["4</w>"], e["2</w>"]]   # 's split, digits single
    assert tk.encode("&amp;") == [e["&</w>"]]                                    # html unescape
    ids = tk.tokenize(["cat", "dog dog dog dog"], context_length=6)
    assert ids.dtype == torch.int64 and ids.shape == (2, 6)
    assert ids[0].tolist() == [tk.sot_id, e["cat</w>"], tk.eot_id, 0, 0, 0]
    assert ids[1].tolist() == [tk.sot_id, e["do"], e["g</w>"], e["do"], e["g</w>"], tk.eot_id]   # truncated, EOT kept
    assert (ids.argmax(-1) == torch.tensor([2, 5])).all()                        # the EOT position clip.py:150 reads
    full = vocab.SimpleTokenizer(merges=[])
    assert full.tokenize("")[0, :2].tolist() == [full.sot_id, full.eot_id]
