"""The reference's OWN experiment INIs for the hot path (tests/{bahdanau,transformer,beamsearch,small}.ini
of /root/reference, the ones `tests/tests_run.sh` trains), unchanged, through this package's
`neuralmonkey-train` entry point - on the CPU over the stand-in operations of tests/cpu_ops.py.
What is exercised is everything but the kernels: the INI grammar with variables and environment
substitution, `class=` resolution against this package, constructor signatures and validation, datasets
with bucketing, vocabularies, the training loop with validation, runners, evaluators, checkpoints.
Only the output locations are redirected (the reference tree is read-only), and `evaluators.TER`
(third-party pyter, absent here as TensorFlow is) is dropped from small.ini's evaluation list.

Skipped when /root/reference is not there (it exists in the build container only)."""
import os
import sys

import pytest
import torch

from tests import cpu_ops

REFERENCE = "/root/reference"
pytestmark = [pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "tests", "data")),
                                 reason="the reference tree is not mounted"),
              pytest.mark.filterwarnings("ignore:Converting a tensor with requires_grad")]

CASES = {
    "bahdanau": ['val_data_no_target.outputs=[("encoded", "{out}/encoded"), ("debugtensors", "{out}/debugtensors")]'],
    "transformer": [],
    "beamsearch": [],
    "small": ['main.evaluation=[("target", $bleu), ("target", evaluators.ChrF3)]'],
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_ini_trains_unchanged(monkeypatch, tmp_path, name):
    from neuralmonkey_b200 import ops, runtime
    from neuralmonkey_b200.trainers.generic_trainer import GenericTrainer
    for op in cpu_ops.STAND_INS:
        monkeypatch.setattr(ops, op, getattr(cpu_ops, op))
    monkeypatch.setattr(runtime, "_device", torch.device("cpu"))
    monkeypatch.setattr(GenericTrainer, "_adam_kernel", cpu_ops.adam_kernel)
    monkeypatch.setenv("NMB200_UNVERIFIED", "1")          # small.ini: Nematus GRU cells, conditional GRU
    monkeypatch.setenv("NEURALMONKEY_STRICT", "1")        # as tests_run.sh: warnings are errors
    monkeypatch.setenv("NM_EXPERIMENT_NAME", "small")     # small.ini reads it from the environment
    monkeypatch.chdir(REFERENCE)                          # the INIs name their data relative to the tree
    out = str(tmp_path / name)
    argv = ["neuralmonkey-train", "tests/{}.ini".format(name), "-s", 'main.output="{}"'.format(out)]
    for change in CASES[name]:
        argv += ["-s", change.format(out=out)]
    monkeypatch.setattr(sys, "argv", argv)
    try:
        from neuralmonkey_b200.train import main
        main()
    finally:
        runtime.reset()
    log_text = open(os.path.join(out, "experiment.log")).read()
    assert "Training finished" in log_text and "Validation (epoch" in log_text
    assert os.path.exists(os.path.join(out, "variables.data.final"))
    if name == "bahdanau":
        assert os.path.exists(os.path.join(out, "encoded.npy"))
    if name == "beamsearch":
        assert "beam_search_score" in log_text
