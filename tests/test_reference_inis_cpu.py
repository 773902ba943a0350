"""The reference's OWN experiment INIs for the hot path (tests/{bahdanau,transformer,beamsearch,small,
factored,post-edit}.ini of /root/reference, the ones `tests/tests_run.sh` trains), unchanged, through this package's
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
    # FactoredEncoder + attention.ScaledDotProdAttention as the RNN decoder's attention object
    "factored": [],
    # two encoders (GRU and LSTM), MultiHeadAttention (3 heads, keys and values from different encoders) +
    # ScaledDotProdAttention on one decoder, the edit-operation pre/postprocessors; pyter's TER dropped
    "post-edit": ['main.evaluation=[("target", <bleu>)]'],
    # an RNN decoder without encoders, word2vec-initialised embeddings, XentRunner + PerplexityEvaluator
    "language-model": [],
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
    if name == "language-model":
        assert "xents/perplexity" in log_text


_FEATURES_INI = """
[main]
name="captioning over pre-extracted feature maps"
tf_manager=<tf_manager>
output="{out}"
overwrite_output_dir=True
batch_size=2
epochs=2
train_dataset=<train_data>
val_dataset=<val_data>
trainer=<trainer>
runners=[<runner>]
postprocess=None
evaluation=[("target", evaluators.BLEU)]
logging_period=1
validation_period=5
random_seed=1234

[tf_manager]
class=tf_manager.TensorFlowManager
num_threads=4
num_sessions=1

[numpy_reader]
class=readers.numpy_reader.from_file_list
prefix="tests/data/flickr30k"
shape=[8, 8, 2048]

[train_data]
class=dataset.load
series=["target", "images"]
data=["tests/data/flickr30k/train.de", ("tests/data/flickr30k/train_images.npz.txt", <numpy_reader>)]

[val_data]
class=dataset.load
series=["target", "images"]
data=["tests/data/flickr30k/val.de", ("tests/data/flickr30k/val_images.npz.txt", <numpy_reader>)]

[imagenet]
class=encoders.numpy_stateful_filler.SpatialFiller
name="imagenet"
input_shape=[8, 8, 2048]
data_id="images"
projection_dim=6
ff_hidden_dim=9

[decoder_vocabulary]
class=vocabulary.from_wordlist
path="tests/data/decoder_vocab.tsv"

[attention]
class=attention.Attention
name="attention"
encoder=<imagenet>
state_size=5

[decoder]
class=decoders.decoder.Decoder
name="decoder"
attentions=[<attention>]
encoders=[<imagenet>]
rnn_size=3
embedding_size=3
dropout_keep_prob=0.5
data_id="target"
max_output_len=3
vocabulary=<decoder_vocabulary>

[trainer]
class=trainers.cross_entropy_trainer.CrossEntropyTrainer
decoders=[<decoder>]
l2_weight=1.0e-8
clip_norm=1.0

[runner]
class=runners.GreedyRunner
decoder=<decoder>
output_series="target"
"""


def test_captioning_over_the_reference_feature_files(monkeypatch, tmp_path):
    """The image half of the reference's tests/flat-multiattention.ini - `readers.numpy_reader.from_file_list`
    over its flickr30k .npz feature maps into a `SpatialFiller` - under the attention decoder of the hot path
    (the FlatMultiAttention wrapper of that INI is outside it)."""
    from neuralmonkey_b200 import ops, runtime
    from neuralmonkey_b200.trainers.generic_trainer import GenericTrainer
    for op in cpu_ops.STAND_INS:
        monkeypatch.setattr(ops, op, getattr(cpu_ops, op))
    monkeypatch.setattr(runtime, "_device", torch.device("cpu"))
    monkeypatch.setattr(GenericTrainer, "_adam_kernel", cpu_ops.adam_kernel)
    monkeypatch.setenv("NEURALMONKEY_STRICT", "1")
    monkeypatch.chdir(REFERENCE)
    out = str(tmp_path / "features")
    ini = tmp_path / "features.ini"
    ini.write_text(_FEATURES_INI.format(out=out))
    monkeypatch.setattr(sys, "argv", ["neuralmonkey-train", str(ini)])
    try:
        from neuralmonkey_b200.train import main
        main()
        names = set(runtime.arena().names) if hasattr(runtime.arena(), "names") else set()
    finally:
        runtime.reset()
    log_text = open(os.path.join(out, "experiment.log")).read()
    assert "Training finished" in log_text and "Validation (epoch" in log_text and "target/BLEU" in log_text
    assert os.path.exists(os.path.join(out, "variables.data.final"))
    saved = torch.load(os.path.join(out, "variables.data.final"), weights_only=False) \
        if not names else None
    keys = names or set(saved.get("variables", saved).keys())
    assert {"imagenet/conv2d/kernel", "imagenet/conv2d_1/kernel"} <= keys
