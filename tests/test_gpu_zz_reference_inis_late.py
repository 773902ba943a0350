"""The reference's tests/factored.ini, tests/post-edit.ini and tests/language-model.ini UNCHANGED through
`neuralmonkey-train` on the GPU (FactoredEncoder + ScaledDotProdAttention; two encoders, MultiHeadAttention +
ScaledDotProdAttention and the edit-operation processors; an RNN decoder without encoders over word2vec
embeddings with the XentRunner / PerplexityEvaluator) - the companion of tests/test_gpu_reference_inis.py for the
INIs that became trainable in the CPU-only part of round 2 (tests/test_reference_inis_cpu.py trains them over the
stand-in operations).  Inputs come from tests/golden/reference_experiments_late.json.

Never run on a GPU yet, hence opt-in (`NMB200_RUN_UNRUN_GPU_TESTS=1`); `bench.py` runs it in a separate
process and records the outcome under `extra_workloads.late_gpu_checks`."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("NMB200_RUN_UNRUN_GPU_TESTS", "0") != "1",
                                 reason="never run on a GPU yet: opt in with NMB200_RUN_UNRUN_GPU_TESTS=1")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUNDLE = os.path.join(ROOT, "tests", "golden", "reference_experiments_late.json")

CASES = {
    "factored": [],
    "post-edit": ['main.evaluation=[("target", <bleu>)]'],        # pyter's TER: third-party, absent
    "language-model": [],
}


def unpack(tree: str) -> None:
    with open(BUNDLE, encoding="utf-8") as handle:
        bundle = json.load(handle)
    for rel, text in bundle["files"].items():
        path = os.path.join(tree, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w", encoding="utf-8") as handle:
            handle.write(text)


def command(name: str, out: str):
    cmd = [sys.executable, os.path.join(ROOT, "bin", "neuralmonkey-train"), "tests/{}.ini".format(name),
           "-s", 'main.output="{}"'.format(out)]
    for change in CASES[name]:
        cmd += ["-s", change]
    return cmd


@pytest.mark.parametrize("name", sorted(CASES))
def test_late_reference_ini_trains_unchanged_on_the_gpu(tmp_path, name):
    tree, out = str(tmp_path / "tree"), str(tmp_path / "out")
    unpack(tree)
    env = dict(os.environ, NEURALMONKEY_STRICT="1", PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    res = subprocess.run(command(name, out), capture_output=True, text=True, timeout=900, cwd=tree, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    log_text = open(os.path.join(out, "experiment.log")).read()
    assert "Training finished" in log_text and "Validation (epoch" in log_text
    assert os.path.exists(os.path.join(out, "variables.data.final"))
    if name == "language-model":
        assert "xents/perplexity" in log_text
