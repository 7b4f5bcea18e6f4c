"""Builds ovmr_b200/clip/bpe_merges.xz — the 48,894 BPE merge rules CLIP's tokenizer uses.

The merge table is third-party DATA (OpenAI CLIP's byte-level BPE vocabulary, MIT licence), not
code; bit-exact token ids are impossible without it.  The reference ships it as
clip/bpe_simple_vocab_16e6.txt.gz and reads lines [1, 48895) (clip/simple_tokenizer.py:66-68); this
script extracts exactly those rules and stores them LZMA-compressed, one "left right" pair per line.

    python tools/build_bpe_table.py [/root/reference/clip/bpe_simple_vocab_16e6.txt.gz]
"""
import gzip
import lzma
import os
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/clip/bpe_simple_vocab_16e6.txt.gz"
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ovmr_b200", "clip", "bpe_merges.xz")
lines = gzip.open(src).read().decode("utf-8").split("\n")
rules = lines[1:49152 - 256 - 2 + 1]
assert len(rules) == 48894 and all(len(r.split()) == 2 for r in rules)
with lzma.open(dst, "wt", encoding="utf-8", preset=9) as f:
    f.write("\n".join(rules))
print(f"wrote {dst}: {len(rules)} merge rules, {os.path.getsize(dst)} bytes")
