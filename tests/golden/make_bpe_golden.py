"""Golden BPE segmentations produced by the reference's own vendored subword-nmt
(/root/reference/lib/subword_nmt/apply_bpe.py), run in this container.
    python tests/golden/make_bpe_golden.py
"""
import io
import json
import os
import sys

sys.path.insert(0, "/root/reference")
from lib.subword_nmt.apply_bpe import BPE, encode  # noqa: E402

MERGES = ["t h", "th e</w>", "i n", "e r", "a n", "r e", "in g</w>", "o n", "e n", "an d</w>", "l o", "lo w",
          "e s", "es t</w>", "n e", "ne w", "w i", "d e", "er </w>", "low est</w>", "t h"]
SENTENCES = [["the", "lowest", "newer", "wider", "and", "thing"], ["a", "", "xyz", "then", "inning"],
             ["lower", "rethink", "onenew"]]

if __name__ == "__main__":
    bpe = BPE(io.StringIO("\n".join(MERGES) + "\n"), "@@")
    out = []
    for sent in SENTENCES:
        seg = []
        for word in sent:
            if not word:
                seg.append(word)
                continue
            pieces = encode(word, bpe.bpe_codes)
            seg.extend(p + "@@" for p in pieces[:-1])
            seg.append(pieces[-1])
        out.append(seg)
    here = os.path.dirname(os.path.abspath(__file__))
    json.dump({"merges": MERGES, "sentences": SENTENCES, "segmented": out},
              open(os.path.join(here, "bpe_golden.json"), "w"), indent=1)
    print(out)
