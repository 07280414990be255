"""Vocabulary front end: label files -> prompt strings -> CLIP BPE token ids -> text bank.

  * read_label_file / prompt_labels / overlapping_mask restate odise/data/build.py:18-71 and the class-overlap rule of
    PoolingCLIPHead.forward (odise/modeling/meta_arch/odise.py:1483-1493); pinned against the reference's own functions and
    label files in tests/test_vocab_cpu.py.
  * SimpleTokenizer restates the CLIP byte-pair tokenizer that `open_clip.tokenize` applies (open-clip-torch==2.0.2,
    setup.py:85, call site clip.py:64): un-vendored third-party code, published algorithm (OpenAI CLIP
    simple_tokenizer.py).  Its merge table (`bpe_simple_vocab_16e6.txt.gz`) ships with open_clip, not with this repo:
    the caller passes its path.  `ftfy.fix_text` (mojibake repair) is not applied — prompts are plain ASCII class names.
"""
import gzip
import html
import re
from functools import lru_cache

import torch

try:                                    # \\p{L} / \\p{N} classes need the `regex` module; ASCII fallback otherwise
    import regex as _re
    _PAT = r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+"
except ImportError:                     # pragma: no cover
    _re = re
    _PAT = r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[a-zA-Z]+|[0-9]|[^\sa-zA-Z0-9]+"

INVALID_NAME = "invalid_class_id"


def read_label_file(path):
    """get_openseg_labels (build.py:18-51) on an explicit file: lines `id:name1,name2,...` -> [[name1, name2, ...], ...]
    (entries named `invalid_class_id` are dropped)."""
    out = []
    with open(path, "r") as f:
        for line in f.read().splitlines():
            _, name = line.split(":")
            if name != INVALID_NAME:
                out.append(name.split(","))
    return out


def prompt_labels(labels, prompt):
    """build.py:54-71."""
    if prompt is None:
        return [list(l) for l in labels]
    fmt = {"a": "a {}", "photo": "a photo of a {}.", "scene": "a photo of a {} in the scene."}
    if prompt not in fmt:
        raise ValueError(f"prompt must be one of {sorted(fmt)} or None")
    return [[fmt[prompt].format(l) for l in syn] for syn in labels]


def overlapping_mask(test_labels, train_labels):
    """odise.py:1483-1493: class k of the test vocabulary counts as seen in training (-> exponent alpha) iff one of its
    synonyms is a synonym of some training class."""
    train = {l for label in train_labels for l in label}
    return [not train.isdisjoint(set(t)) for t in test_labels]


@lru_cache()
def bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def _pairs(word):
    return {(a, b) for a, b in zip(word[:-1], word[1:])}


class SimpleTokenizer:
    SOT, EOT = "<|startoftext|>", "<|endoftext|>"

    def __init__(self, bpe_path=None, merges=None):
        """bpe_path: CLIP's bpe_simple_vocab_16e6.txt(.gz); or `merges`: an explicit list of "left right" merge rules."""
        self.byte_encoder = bytes_to_unicode()
        if merges is None:
            opener = gzip.open if str(bpe_path).endswith(".gz") else open
            with opener(bpe_path, "rt", encoding="utf-8") as f:
                lines = f.read().split("\n")
            merges = lines[1:49152 - 256 - 2 + 1]
        merges = [tuple(m.split()) for m in merges if m]
        vocab = list(self.byte_encoder.values())
        vocab = vocab + [v + "</w>" for v in vocab] + ["".join(m) for m in merges] + [self.SOT, self.EOT]
        self.encoder = {t: i for i, t in enumerate(vocab)}
        self.bpe_ranks = {m: i for i, m in enumerate(merges)}
        self.cache = {self.SOT: self.SOT, self.EOT: self.EOT}
        self.pat = _re.compile(_PAT, _re.IGNORECASE)
        self.sot_id, self.eot_id = self.encoder[self.SOT], self.encoder[self.EOT]

    def bpe(self, token):
        if token in self.cache:
            return self.cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        pairs = _pairs(word)
        if not pairs:
            return token + "</w>"
        while True:
            bigram = min(pairs, key=lambda p: self.bpe_ranks.get(p, float("inf")))
            if bigram not in self.bpe_ranks:
                break
            first, second = bigram
            new, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == first and word[i + 1] == second:
                    new.append(first + second)
                    i += 2
                else:
                    new.append(word[i])
                    i += 1
            word = tuple(new)
            if len(word) == 1:
                break
            pairs = _pairs(word)
        out = " ".join(word)
        self.cache[token] = out
        return out

    def encode(self, text):
        text = re.sub(r"\s+", " ", html.unescape(html.unescape(text)).strip()).strip().lower()
        ids = []
        for tok in self.pat.findall(text):
            tok = "".join(self.byte_encoder[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self.bpe(tok).split(" "))
        return ids

    def tokenize(self, texts, context_length=77):
        """open_clip.tokenize: [SOT] + ids + [EOT], truncated to context_length with EOT kept last, zero padded -> int64."""
        if isinstance(texts, str):
            texts = [texts]
        out = torch.zeros(len(texts), context_length, dtype=torch.int64)
        for i, t in enumerate(texts):
            ids = [self.sot_id] + self.encode(t) + [self.eot_id]
            if len(ids) > context_length:
                ids = ids[:context_length]
                ids[-1] = self.eot_id
            out[i, :len(ids)] = torch.tensor(ids)
        return out


def vocabulary_prompts(test_labels, category_prompt=None, clip_prompt="photo"):
    """The two prompt sets of one test vocabulary: CategoryEmbed scores against prompt_labels(labels, None) = the raw
    class names (odise.py:1225, :1302; mask_generator_with_label.py builds CategoryEmbed without a prompt), PoolingCLIPHead
    against prompt_labels(labels, "photo") (odise.py:1428, :1475).  The caption model (CLIPOpenClassEmbed /
    WordEmbed, odise.py:1026) uses "photo" for both.  -> (flat category prompts, flat clip prompts, group sizes)"""
    cat, clip = prompt_labels(test_labels, category_prompt), prompt_labels(test_labels, clip_prompt)
    return [p for syn in cat for p in syn], [p for syn in clip for p in syn], [len(s) for s in cat]


def build_vocabulary(engine, tokenizer, key, test_labels, train_labels=None, thing_ids=None, category_prompt=None,
                     clip_prompt="photo"):
    """CategoryEmbed.forward eval branch + PoolingCLIPHead's label handling (odise.py:1298-1307, :1476-1497): class synonym
    lists -> the two prompt sets -> token ids -> two CLIP text banks on the device -> engine vocabulary.
    train_labels (PoolingCLIPHead.train_labels, default in the reference: the COCO panoptic prompt-engineered label file)
    decides which classes take exponent alpha; it is required whenever the engine has a MaskCLIP head."""
    cat, clip, sizes = vocabulary_prompts(test_labels, category_prompt, clip_prompt)
    ov = overlapping_mask(test_labels, train_labels) if train_labels is not None else None
    clip_ids = None if clip == cat else tokenizer.tokenize(clip)
    return engine.set_vocabulary_from_tokens(key, tokenizer.tokenize(cat), sizes, thing_ids=thing_ids, overlapping=ov,
                                             clip_token_ids=clip_ids)
