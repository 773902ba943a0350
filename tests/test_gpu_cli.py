"""End-to-end drop-in check: an INI written in Neural Monkey's own format (modelled on the
reference's tests/bahdanau.ini: bucketed datasets, wordlist vocabularies, SentenceEncoder +
Attention + maxout Decoder with dropout and <unk> suppression, MultitaskTrainer over three
CrossEntropyTrainers, Greedy/Representation/Tensor runners) is trained with
bin/neuralmonkey-train and decoded with bin/neuralmonkey-run."""
import json
import os
import random
import subprocess
import sys

import pytest

from tests.helpers import training_log_values

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

INI = """
[main]
name="toy reversal bahdanau style"
tf_manager=<tf_manager>
output="{out}"
overwrite_output_dir=True
batch_size=16
epochs={epochs}
train_dataset=<train_data>
val_dataset=[<val_data>,<val_data>]
trainer=[<mt_trainer>, <greedy_trainer>]
runners=[<runner>, <representation_runner>, <debug_runner>]
postprocess=None
evaluation=[("target", evaluators.ROUGE_L), ("target", evaluators.SacreBLEU)]
logging_period=20
validation_period=40
test_datasets=[<val_data_no_target>]

[tf_manager]
class=tf_manager.TensorFlowManager
num_threads=4
num_sessions=1

[batching]
class=dataset.BatchingScheme
bucket_boundaries=[4, 6, 8]
bucket_batch_sizes=[20, 15, 10, 5]

[train_data]
class=dataset.load
series=["source", "target"]
data=["{data}/train.src", "{data}/train.tgt"]
batching=<batching>

[val_data]
class=dataset.load
series=["source", "target"]
data=["{data}/val.src", "{data}/val.tgt"]
batching=<batching>

[val_data_no_target]
class=dataset.load
series=["source"]
data=["{data}/val.src"]
outputs=[("target", "{out}/val.out"), ("encoded", "{out}/encoded"), ("debugtensors", "{out}/debugtensors")]
batching=<batching>

[encoder_vocabulary]
class=vocabulary.from_wordlist
path="{data}/src_vocab.tsv"

[encoder]
class=encoders.recurrent.SentenceEncoder
name="sentence_encoder"
rnn_size=7
max_input_len=10
embedding_size=11
data_id="source"
vocabulary=<encoder_vocabulary>

[attention]
class=attention.Attention
name="attention_sentence_encoder"
encoder=<encoder>

[decoder_vocabulary]
class=vocabulary.from_wordlist
path="{data}/tgt_vocab.tsv"

[decoder]
class=decoders.decoder.Decoder
name="bahdanau_decoder"
encoders=[<encoder>]
rnn_size=8
embedding_size=9
attentions=[<attention>]
output_projection=<dec_maxout_output>
dropout_keep_prob=0.5
data_id="target"
max_output_len=10
vocabulary=<decoder_vocabulary>
supress_unk=True

[dec_maxout_output]
class=decoders.output_projection.maxout_output
maxout_size=9

[trainer1]
class=trainers.cross_entropy_trainer.CrossEntropyTrainer
decoders=[<decoder>]
l2_weight=1.0e-8
clip_norm=1.0
optimizer=<adam>

[trainer2]
class=trainers.cross_entropy_trainer.CrossEntropyTrainer
decoders=[<decoder>]
optimizer=<adam>

[greedy_trainer]
class=trainers.cross_entropy_trainer.CrossEntropyTrainer
decoders=[<decoder>]
clip_norm=10
l1_weight=0.0001
optimizer=<adam>

[adam]
class=tf.train.AdamOptimizer
learning_rate=1.0e-2

[mt_trainer]
class=trainers.multitask_trainer.MultitaskTrainer
trainers=[<trainer1>, <trainer1>, <trainer2>]

[runner]
class=runners.GreedyRunner
output_series="target"
decoder=<decoder>

[representation_runner]
class=runners.tensor_runner.RepresentationRunner
encoder=<encoder>
output_series="encoded"

[debug_runner]
class=runners.tensor_runner.TensorRunner
modelparts=[<encoder>, <encoder>, <decoder>]
tensors=["output", "temporal_states", "runtime_logits"]
batch_dims=[0, 0, 1]
tensors_by_name=[]
batch_dims_by_name=[]
output_series="debugtensors"
"""


def _write_data(path):
    rng = random.Random(0)
    words = ["w{}".format(i) for i in range(20)]
    os.makedirs(path, exist_ok=True)

    def vocab(name, ws):
        with open(os.path.join(path, name), "w") as f:
            f.write("word\tcount\n<pad>\t0\n<s>\t0\n</s>\t0\n<unk>\t0\n")
            for w in ws:
                f.write("{}\t1\n".format(w))

    vocab("src_vocab.tsv", words)
    vocab("tgt_vocab.tsv", [w.upper() for w in words])
    for split, n in (("train", 300), ("val", 30)):
        with open(os.path.join(path, split + ".src"), "w") as fs, \
                open(os.path.join(path, split + ".tgt"), "w") as ft:
            for _ in range(n):
                sent = [rng.choice(words) for _ in range(rng.randint(1, 8))]
                fs.write(" ".join(sent) + "\n")
                ft.write(" ".join(w.upper() for w in reversed(sent)) + "\n")


def _run(cmd, **kw):
    env = dict(os.environ, NEURALMONKEY_STRICT="1")
    return subprocess.run([sys.executable] + cmd, capture_output=True, text=True, timeout=900,
                          cwd=ROOT, env=env, **kw)


def test_train_and_run_entry_points(tmp_path):
    data, out = str(tmp_path / "data"), str(tmp_path / "out")
    _write_data(data)
    ini = tmp_path / "toy.ini"
    ini.write_text(INI.format(out=out, data=data, epochs=3))
    res = _run(["bin/neuralmonkey-train", str(ini)])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    for name in ("experiment.ini", "original.ini", "experiment.log", "variables.data",
                 "variables.data.best", "variables.data.final", "val.out"):
        assert os.path.exists(os.path.join(out, name)), name
    log_text = open(os.path.join(out, "experiment.log")).read()
    assert "Validation (epoch" in log_text and "target/BLEU" in log_text
    # the loss must have gone down during the three epochs
    losses = training_log_values(log_text, "target/train_xent")
    assert len(losses) >= 3 and losses[-1] < losses[0], losses
    assert len(open(os.path.join(out, "val.out")).read().splitlines()) == 30

    # neuralmonkey-run with a datasets INI, as tests/tests_run.sh does
    run_ini = tmp_path / "run.ini"
    run_ini.write_text("""
[main]
test_datasets=[<val_data>]

[batching]
class=dataset.BatchingScheme
batch_size=7

[val_data]
class=dataset.load
series=["source", "target"]
data=["{data}/val.src", "{data}/val.tgt"]
outputs=[("target", "{out}/run.out")]
batching=<batching>
""".format(data=data, out=out))
    res = _run(["bin/neuralmonkey-run", str(ini), str(run_ini), "--json", str(tmp_path / "res.json")])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    results = json.load(open(tmp_path / "res.json"))
    assert "target/BLEU" in results[0] and "target/runtime_xent" in results[0]
    assert len(open(os.path.join(out, "run.out")).read().splitlines()) == 30


# Modelled on the reference's tests/transformer.ini + tests/beamsearch.ini: EmbeddedSequence +
# TransformerEncoder/Decoder with dropout, DelayedUpdateTrainer, LazyAdam with the Noam schedule, a
# char-level preprocessed series, BeamSearchDecoder + beam_search_runner_range, BLEU on rank 1.
TRANSFORMER_INI = """
[main]
name="toy reversal transformer with beam search"
tf_manager=<tf_manager>
output="{out}"
overwrite_output_dir=True
batch_size=16
epochs={epochs}
train_dataset=<train_data>
val_dataset=<val_data>
trainer=<trainer>
runners=<bs_runners>
postprocess=None
evaluation=[("target_beam.rank001", "target", evaluators.BLEU)]
logging_period=10
validation_period=20
random_seed=1234

[tf_manager]
class=tf_manager.TensorFlowManager
num_threads=4
num_sessions=1
save_n_best=2

[train_data]
class=dataset.load
series=["source", "target", "source_chars"]
data=["{data}/train.src", "{data}/train.tgt", (processors.helpers.preprocess_char_based, "source")]
buffer_size=32

[val_data]
class=dataset.load
series=["source", "target", "source_chars"]
data=["{data}/val.src", "{data}/val.tgt", (processors.helpers.preprocess_char_based, "source")]

[encoder_vocabulary]
class=vocabulary.from_wordlist
path="{data}/src_vocab.tsv"

[inpseq]
class=model.sequence.EmbeddedSequence
name="input"
embedding_size=12
max_length=9
data_id="source"
vocabulary=<encoder_vocabulary>

[encoder]
class=encoders.transformer.TransformerEncoder
name="transformer_encoder"
input_sequence=<inpseq>
ff_hidden_size=20
depth=2
n_heads=3
dropout_keep_prob=0.9

[decoder_vocabulary]
class=vocabulary.from_wordlist
path="{data}/tgt_vocab.tsv"

[decoder]
class=decoders.transformer.TransformerDecoder
name="decoder"
encoders=[<encoder>]
dropout_keep_prob=0.9
data_id="target"
max_output_len=9
vocabulary=<decoder_vocabulary>
embedding_size=12
ff_hidden_size=20
depth=2
n_heads_self=3
n_heads_enc=2

[trainer]
class=trainers.delayed_update_trainer.DelayedUpdateTrainer
batches_per_update=2
l2_weight=1.0e-8
clip_norm=1.0
objectives=[<obj>]
optimizer=<lazyadam_g>

[obj]
class=trainers.cross_entropy_trainer.CostObjective
decoder=<decoder>

[decayed_lr]
class=functions.noam_decay
learning_rate=0.5
model_dimension=12
warmup_steps=20

[lazyadam_g]
class=tf.contrib.opt.LazyAdamOptimizer
beta1=0.9
beta2=0.98
epsilon=1.0e-9
learning_rate=<decayed_lr>

[bs_decoder]
class=decoders.beam_search_decoder.BeamSearchDecoder
name="beam_search_decoder"
parent_decoder=<decoder>
length_normalization=0.6
max_steps=10
beam_size=3

[bs_runners]
class=runners.beam_search_runner_range
output_series="target_beam"
decoder=<bs_decoder>
max_rank=2
"""


def test_transformer_beam_search_experiment(tmp_path):
    data, out = str(tmp_path / "data"), str(tmp_path / "out")
    _write_data(data)
    ini = tmp_path / "transformer.ini"
    ini.write_text(TRANSFORMER_INI.format(out=out, data=data, epochs=4))
    res = _run(["bin/neuralmonkey-train", str(ini)])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    log_text = open(os.path.join(out, "experiment.log")).read()
    assert "Validation (epoch" in log_text and "target_beam.rank001/BLEU" in log_text
    train_values = training_log_values(log_text, "target_beam.rank001/beam_search_score")
    assert len(train_values) >= 3, log_text[-2000:]
    assert "beam_search_score" in log_text
    assert os.path.exists(os.path.join(out, "variables.data.best"))
