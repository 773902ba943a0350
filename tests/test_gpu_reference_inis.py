"""The reference's OWN experiment INIs for the hot path - tests/bahdanau.ini, tests/transformer.ini,
tests/beamsearch.ini, the ones its tests/tests_run.sh trains - UNCHANGED through this package's
`neuralmonkey-train` entry point on the GPU: INI grammar, `class=` resolution, constructors, bucketed
datasets, the CUDA kernels behind every model part, trainers (MultitaskTrainer over two CrossEntropyTrainers;
DelayedUpdateTrainer with LazyAdam + Noam), greedy and beam-search runners, evaluators, checkpoints.

The INIs and the toy corpora come from tests/golden/reference_experiments.json (a bundle of the reference's
files made by tests/golden/make_reference_bundle.py: the GPU box has no /root/reference); only the output
directory is redirected.  `NEURALMONKEY_STRICT=1` as in tests_run.sh:5: warnings are errors."""
import json
import os
import subprocess
import sys

import pytest

from tests.helpers import training_log_values

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUNDLE = os.path.join(ROOT, "tests", "golden", "reference_experiments.json")

CASES = {
    "bahdanau": ['val_data_no_target.outputs=[("encoded", "{out}/encoded"), ("debugtensors", "{out}/debugtensors")]'],
    "transformer": [],
    "beamsearch": [],
}


def _unpack(tree: str) -> None:
    with open(BUNDLE, encoding="utf-8") as handle:
        bundle = json.load(handle)
    for rel, text in bundle["files"].items():
        path = os.path.join(tree, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w", encoding="utf-8") as handle:
            handle.write(text)


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_ini_trains_unchanged_on_the_gpu(tmp_path, name):
    tree, out = str(tmp_path / "tree"), str(tmp_path / "out")
    _unpack(tree)
    cmd = [sys.executable, os.path.join(ROOT, "bin", "neuralmonkey-train"), "tests/{}.ini".format(name),
           "-s", 'main.output="{}"'.format(out)]
    for change in CASES[name]:
        cmd += ["-s", change.format(out=out)]
    env = dict(os.environ, NEURALMONKEY_STRICT="1", PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=tree, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    log_text = open(os.path.join(out, "experiment.log")).read()
    assert "Training finished" in log_text and "Validation (epoch" in log_text
    assert os.path.exists(os.path.join(out, "variables.data.final"))
    if name == "bahdanau":
        assert os.path.exists(os.path.join(out, "encoded.npy"))
        losses = training_log_values(log_text, "target/train_xent")
        # finite; NOT small: the INI sets supress_unk=True and the toy references contain <unk>, whose logit
        # carries -1e9 in training as well (autoregressive.py:454-457) - the reference logs ~3e8 here too
        assert losses and all(l == l and abs(l) != float("inf") for l in losses), losses
    if name == "beamsearch":
        assert "beam_search_score" in log_text
