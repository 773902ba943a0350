"""Bundle the INPUTS of the reference's own hot-path experiments - tests/{bahdanau,transformer,beamsearch}.ini
and the toy corpora / vocabularies they name - into tests/golden/reference_experiments.json, so that the
`-m gpu` tests can train them UNCHANGED on a box that has no /root/reference (tests/test_gpu_reference_inis.py
unpacks the bundle into a scratch tree and runs `neuralmonkey-train tests/<name>.ini` from its root).
These are data fixtures (configurations and corpora), the same role tests/golden/*.npz play for tensors.

    python tests/golden/make_reference_bundle.py        # needs /root/reference
"""
import json
import os

REFERENCE = "/root/reference"
FILES = ["tests/bahdanau.ini", "tests/transformer.ini", "tests/beamsearch.ini",
         "tests/data/train.tc.en", "tests/data/train.tc.de", "tests/data/val.tc.en", "tests/data/val.tc.de",
         "tests/data/encoder_vocab.tsv", "tests/data/decoder_vocab.tsv"]


def main() -> None:
    bundle = {"source": "ufal/neuralmonkey @ 8b1465270f6bb28d5417a85cec492f7179036ede", "files": {}}
    for rel in FILES:
        with open(os.path.join(REFERENCE, rel), encoding="utf-8") as handle:
            bundle["files"][rel] = handle.read()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_experiments.json")
    with open(out, "w", encoding="utf-8") as handle:
        json.dump(bundle, handle, ensure_ascii=False, indent=0)
    print("wrote", out, sum(len(v) for v in bundle["files"].values()), "characters")


if __name__ == "__main__":
    main()
