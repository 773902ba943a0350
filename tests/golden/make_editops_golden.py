"""Golden vectors for processors/editops.py, produced by the reference's own module run in this container
(it imports numpy only): seeded random word sequences over a small alphabet (many ties between equally
cheap scripts), the sentences of tests/data/postedit, and edge cases (empty sides, scripts longer / shorter
than the source).
    python tests/golden/make_editops_golden.py
"""
import importlib.util
import json
import os
import random

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location(
        "ref_editops", os.path.join(REF, "neuralmonkey", "processors", "editops.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = random.Random(20260923)
    pairs = [([], []), ([], ["a"]), (["a"], []), (["a"], ["a"]), (["a", "b"], ["b", "a"]),
             (["a", "a", "a"], ["a", "a"]), (["a", "a"], ["a", "a", "a"]), (["x", "y", "z"], ["p", "q"])]
    for _ in range(120):
        alphabet = "abc" if rng.random() < 0.5 else "abcdefgh"
        pairs.append(([rng.choice(alphabet) for _ in range(rng.randint(0, 9))],
                      [rng.choice(alphabet) for _ in range(rng.randint(0, 9))]))
    data = os.path.join(REF, "tests", "data", "postedit")
    with open(os.path.join(data, "train.mt"), encoding="utf-8") as f_mt, \
            open(os.path.join(data, "train.pe"), encoding="utf-8") as f_pe:
        for mt, pe in list(zip(f_mt, f_pe))[:40]:
            pairs.append((mt.split(), pe.split()))
    convert = [{"source": s, "target": t, "edits": ref.convert_to_edits(s, t)} for s, t in pairs]
    apply = []
    for case in convert:
        ops = case["edits"]
        variants = [ops, ops[:len(ops) // 2], ops + [ref.KEEP, "extra", ref.DELETE], [ref.KEEP] * 12, []]
        for script in variants:
            apply.append({"source": case["source"], "edits": script,
                          "result": ref.reconstruct(case["source"], script)})
    with open(os.path.join(HERE, "editops_golden.json"), "w", encoding="utf-8") as out:
        json.dump({"convert": convert, "reconstruct": apply}, out, ensure_ascii=False)
    print(len(convert), "conversions,", len(apply), "applications")


if __name__ == "__main__":
    main()
