"""Byte-level BPE tokenizer with CLIP's vocabulary (49,408 ids; SOT=49406, EOT=49407).

Same interface as the reference's clip/simple_tokenizer.py (`SimpleTokenizer.encode / decode`,
`.encoder`, `.decoder`) and token-for-token identical output (pinned by tests/golden/tokenizer.npz),
written from the published algorithm: text -> cleaned lower-case -> regex pre-tokens -> bytes mapped
to printable code points -> greedy lowest-rank pair merging.  The merge table is data
(`bpe_merges.xz`, see tools/build_bpe_table.py).  Host-side only; not a GPU kernel.
"""
import html
import lzma
import os
from functools import lru_cache
from typing import Dict, List, Tuple

import regex as re

try:  # optional: only changes behaviour for mojibake input; ASCII class names are unaffected
    import ftfy

    _fix_text = ftfy.fix_text
except Exception:  # pragma: no cover - ftfy is not installed in the build image
    _fix_text = lambda s: s  # noqa: E731

_END = "</w>"
_SOT, _EOT = "<|startoftext|>", "<|endoftext|>"


@lru_cache()
def default_bpe() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "bpe_merges.xz")


@lru_cache()
def bytes_to_unicode() -> Dict[int, str]:
    """Reversible byte -> printable unicode map: printable latin-1 bytes map to themselves, the other
    68 bytes are moved to code points 256.. so that no token contains whitespace/control chars."""
    keep = [b for b in range(256) if 33 <= b <= 126 or 161 <= b <= 172 or 174 <= b <= 255]
    # the reference enumerates the kept bytes first (in that order) and the rest afterwards
    table, extra = {}, 0
    for b in keep:
        table[b] = chr(b)
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + extra)
            extra += 1
    # dict order = vocabulary order: kept bytes ascending, then remapped bytes ascending
    return table


def basic_clean(text: str) -> str:
    return html.unescape(html.unescape(_fix_text(text))).strip()


def whitespace_clean(text: str) -> str:
    return re.sub(r"\s+", " ", text).strip()


class SimpleTokenizer(object):
    def __init__(self, bpe_path: str = default_bpe()):
        self.byte_encoder = bytes_to_unicode()
        self.byte_decoder = {v: k for k, v in self.byte_encoder.items()}
        with lzma.open(bpe_path, "rt", encoding="utf-8") as f:
            rules: List[Tuple[str, str]] = [tuple(line.split()) for line in f.read().split("\n")]
        symbols = list(self.byte_encoder.values())
        vocab = symbols + [s + _END for s in symbols] + [a + b for a, b in rules] + [_SOT, _EOT]
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.decoder = {i: tok for tok, i in self.encoder.items()}
        self.bpe_ranks = {pair: rank for rank, pair in enumerate(rules)}
        self.cache = {_SOT: _SOT, _EOT: _EOT}
        self.pat = re.compile(
            r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""",
            re.IGNORECASE)

    def bpe(self, token: str) -> str:
        """Merged form of one pre-token as a space-separated symbol string."""
        hit = self.cache.get(token)
        if hit is not None:
            return hit
        word = list(token[:-1]) + [token[-1] + _END]
        inf = len(self.bpe_ranks)
        while len(word) > 1:
            # lowest-rank adjacent pair present in the word
            best_rank, best = inf, None
            for a, b in zip(word, word[1:]):
                r = self.bpe_ranks.get((a, b), inf)
                if r < best_rank:
                    best_rank, best = r, (a, b)
            if best is None:
                break
            a, b = best
            merged, i, n = [], 0, len(word)
            while i < n:  # merge every non-overlapping occurrence, left to right
                if i + 1 < n and word[i] == a and word[i + 1] == b:
                    merged.append(a + b)
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = merged
        out = " ".join(word)
        self.cache[token] = out
        return out

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        text = whitespace_clean(basic_clean(text)).lower()
        for tok in re.findall(self.pat, text):
            tok = "".join(self.byte_encoder[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder[s] for s in self.bpe(tok).split(" "))
        return ids

    def decode(self, tokens) -> str:
        text = "".join(self.decoder[int(t)] for t in tokens)
        return bytearray(self.byte_decoder[c] for c in text).decode("utf-8", errors="replace").replace(_END, " ")
