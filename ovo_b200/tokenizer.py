"""CLIP byte-level BPE tokenizer (host side of Q1).

The reference tokenises with `core.vision_encoder.tokenizer.SimpleTokenizer` (thirdParty/perception_models/
core/vision_encoder/tokenizer.py:132-278), i.e. OpenAI CLIP's published BPE: lower-case, whitespace clean,
regex pre-tokenisation, byte->unicode mapping, greedy lowest-rank merges, <start_of_text>/<end_of_text>
around the ids, zero padding to the context length, truncation keeping EOT last.

The merge table is a data asset (`bpe_simple_vocab_16e6.txt.gz`) that ships with the reference, not with this
repo.  `find_vocab()` locates it through $OVO_B200_BPE_VOCAB or next to an importable
`core.vision_encoder` package (the drop-in runs inside the reference's environment).
Known answer: "a chair" -> [49406, 320, 4269, 49407, 0, ...]."""
import gzip
import html
import os
from functools import lru_cache

import numpy as np
import regex as re

VOCAB_FILE = "bpe_simple_vocab_16e6.txt.gz"


def find_vocab() -> str | None:
    cands = [os.environ.get("OVO_B200_BPE_VOCAB")]
    try:
        import core.vision_encoder as ve          # only importable inside the reference's environment
        cands.append(os.path.join(os.path.dirname(ve.__file__), VOCAB_FILE))
    except Exception:
        pass
    cands.append("/root/reference/thirdParty/perception_models/core/vision_encoder/" + VOCAB_FILE)
    for c in cands:
        if c and os.path.exists(c):
            return c
    return None


@lru_cache()
def _byte_unicode():
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    chars = keep[:]
    extra = 0
    for b in range(256):
        if b not in keep:
            keep.append(b)
            chars.append(256 + extra)
            extra += 1
    return dict(zip(keep, (chr(c) for c in chars)))


class BPETokenizer:
    def __init__(self, vocab_path: str | None = None, context_length: int = 32):
        vocab_path = vocab_path or find_vocab()
        if vocab_path is None:
            raise FileNotFoundError(f"{VOCAB_FILE} not found: set OVO_B200_BPE_VOCAB or pass token ids directly")
        self.context_length = context_length
        self.b2u = _byte_unicode()
        lines = gzip.open(vocab_path).read().decode("utf-8").split("\n")
        merges = [tuple(m.split()) for m in lines[1: 49152 - 256 - 2 + 1]]
        vocab = list(self.b2u.values())
        vocab = vocab + [v + "</w>" for v in vocab] + ["".join(m) for m in merges]
        vocab += ["<start_of_text>", "<end_of_text>"]
        self.encoder = {t: i for i, t in enumerate(vocab)}
        self.ranks = {m: i for i, m in enumerate(merges)}
        self.sot, self.eot = self.encoder["<start_of_text>"], self.encoder["<end_of_text>"]
        self.cache = {}
        self.pat = re.compile(r"""<start_of_text>|<end_of_text>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""",
                              re.IGNORECASE)

    def _bpe(self, token: str):
        if token in self.cache:
            return self.cache[token]
        word = list(token[:-1]) + [token[-1] + "</w>"]
        while len(word) > 1:
            rank, i0 = min((self.ranks.get((a, b), 1 << 30), i) for i, (a, b) in enumerate(zip(word, word[1:])))
            if rank == 1 << 30:
                break
            first, second = word[i0], word[i0 + 1]
            out, i = [], 0
            while i < len(word):                        # merge every occurrence of the best pair, left to right
                if i < len(word) - 1 and word[i] == first and word[i + 1] == second:
                    out.append(first + second)
                    i += 2
                else:
                    out.append(word[i])
                    i += 1
            word = out
        self.cache[token] = word
        return word

    def encode(self, text: str):
        text = html.unescape(html.unescape(text)).strip()
        text = re.sub(r"\s+", " ", text).strip().lower()
        ids = []
        for tok in re.findall(self.pat, text):
            tok = "".join(self.b2u[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self._bpe(tok))
        return ids

    def __call__(self, texts, context_length: int | None = None):
        """str | list[str] -> int64 torch tensor [n, context_length] (same contract as SimpleTokenizer.__call__)."""
        import torch
        if isinstance(texts, str):
            texts = [texts]
        L = context_length or self.context_length
        out = np.zeros((len(texts), L), np.int64)
        for i, t in enumerate(texts):
            ids = [self.sot] + self.encode(t) + [self.eot]
            if len(ids) > L:
                ids = ids[:L]
                ids[-1] = self.eot
            out[i, : len(ids)] = ids
        return torch.from_numpy(out)
