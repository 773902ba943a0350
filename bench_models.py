"""The benchmark models, built the way a user builds them: from an INI through the package's own
configuration builder (`neuralmonkey_b200.config`), with the model sections of the reference's configs
(examples/translation.ini:90-125, tests/transformer.ini:39-102, tests/beamsearch.ini:92-104,
tests/captioning.ini:46-74) at the perf shapes of BASELINE.json / SURVEY.md 8.  Only the data side is
synthetic: a generated word list of the requested size and id tensors fed through `feed_ids`.
"""
import os
import tempfile
from argparse import Namespace

import torch

SEED = 2574600          # the reference's default random_seed (experiment.py:461)

ENDE_INI = """
[main]
encoder=<encoder>
attention=<attention>
decoder=<decoder>
trainer=<trainer>
runner=<runner>
bs_decoder=<bs_decoder>

[shared_vocabulary]
class=vocabulary.from_wordlist
path="{vocab}"
contains_header=False
contains_frequencies=False

; examples/translation.ini:90-125, batch size and vocabulary per BASELINE.json
[encoder]
class=encoders.SentenceEncoder
name="sentence_encoder"
rnn_size=300
max_input_len=50
embedding_size=300
dropout_keep_prob=1.0
data_id="source_bpe"
vocabulary=<shared_vocabulary>

[attention]
class=attention.Attention
name="attention_sentence_encoder"
encoder=<encoder>

[decoder]
class=decoders.Decoder
name="decoder"
encoders=[<encoder>]
rnn_size=300
embedding_size=300
attentions=[<attention>]
dropout_keep_prob=1.0
data_id="target_bpe"
vocabulary=<shared_vocabulary>
max_output_len=50

[trainer]
class=trainers.cross_entropy_trainer.CrossEntropyTrainer
decoders=[<decoder>]
l2_weight=1.0e-8
clip_norm=1.0
use_cuda_graph={graph}

[runner]
class=runners.runner.GreedyRunner
decoder=<decoder>
output_series="target"

; tests/beamsearch.ini:92-98 around the RNN decoder, at the perf shape (beam 8, 128 steps)
[bs_decoder]
class=decoders.beam_search_decoder.BeamSearchDecoder
name="bs_decoder"
parent_decoder=<decoder>
beam_size=8
max_steps={beam_steps}
length_normalization=0.6
"""

TRANSFORMER_INI = """
[main]
input_sequence=<input_sequence>
encoder=<encoder>
decoder=<decoder>
trainer=<trainer>
runner=<runner>
bs_decoder=<bs_decoder>

[shared_vocabulary]
class=vocabulary.from_wordlist
path="{vocab}"
contains_header=False
contains_frequencies=False

; tests/transformer.ini:39-72 at the perf shape: L=6, d=512, h=8, F=2048
[input_sequence]
class=model.sequence.EmbeddedSequence
name="input_sequence"
vocabulary=<shared_vocabulary>
data_id="source"
embedding_size=512
scale_embeddings_by_depth=True
max_length={max_len}

[encoder]
class=encoders.transformer.TransformerEncoder
name="encoder"
input_sequence=<input_sequence>
ff_hidden_size=2048
depth=6
n_heads=8
dropout_keep_prob={keep}
attention_dropout_keep_prob={keep}

[decoder]
class=decoders.transformer.TransformerDecoder
name="decoder"
encoders=[<encoder>]
vocabulary=<shared_vocabulary>
data_id="target"
ff_hidden_size=2048
n_heads_self=8
n_heads_enc=8
depth=6
max_output_len={max_len}
dropout_keep_prob={keep}
attention_dropout_keep_prob={keep}
embedding_size=512
tie_embeddings={tie}

; tests/transformer.ini:79-102: LazyAdam beta2 0.98 eps 1e-9, Noam schedule
[trainer]
class=trainers.cross_entropy_trainer.CrossEntropyTrainer
decoders=[<decoder>]
optimizer=<adam>
use_cuda_graph={graph}

[adam]
class=tf.contrib.opt.LazyAdamOptimizer
beta1=0.9
beta2=0.98
epsilon=1.0e-9
learning_rate=<decayed_learning_rate>

[decayed_learning_rate]
class=functions.noam_decay
learning_rate=0.2
model_dimension=512
warmup_steps=4000

[runner]
class=runners.runner.GreedyRunner
decoder=<decoder>
output_series="target"

; tests/beamsearch.ini:92-98 at the perf shape
[bs_decoder]
class=decoders.beam_search_decoder.BeamSearchDecoder
name="bs_decoder"
parent_decoder=<decoder>
beam_size=8
max_steps={beam_steps}
length_normalization=0.6
"""

CAPTIONING_INI = """
[main]
encoder=<imagenet>
attention=<attention>
decoder=<decoder>
trainer=<trainer>
runner=<runner>

[decoder_vocabulary]
class=vocabulary.from_wordlist
path="{vocab}"
contains_header=False
contains_frequencies=False

; tests/captioning.ini:46-74: VGG-16 conv5_3 maps, Bahdanau decoder (scaled variant of SURVEY.md 8(d))
[imagenet]
class=encoders.imagenet_encoder.ImageNet
name="imagenet_vgg"
data_id="images"
network_type="vgg_16"
spatial_layer="vgg_16/conv5/conv5_3"

[attention]
class=attention.Attention
name="attention"
encoder=<imagenet>
state_size=512

[decoder]
class=decoders.decoder.Decoder
name="decoder"
encoders=[<imagenet>]
attentions=[<attention>]
rnn_size=512
embedding_size=512
max_output_len={max_len}
vocabulary=<decoder_vocabulary>
data_id="target"

[trainer]
class=trainers.cross_entropy_trainer.CrossEntropyTrainer
decoders=[<decoder>]
use_cuda_graph={graph}

[runner]
class=runners.runner.GreedyRunner
decoder=<decoder>
output_series="target"
"""

_TMP = []


def _wordlist(vocab_size: int) -> str:
    """A synthetic word list: the four special tokens, then vocab_size - 4 words."""
    tmp = tempfile.NamedTemporaryFile("w", suffix=".vocab", delete=False)
    tmp.write("<pad>\n<s>\n</s>\n<unk>\n")
    for i in range(vocab_size - 4):
        tmp.write("w{}\n".format(i))
    tmp.close()
    _TMP.append(tmp.name)
    return tmp.name


def build(ini_text: str, fields, **fmt) -> Namespace:
    """INI text -> built objects (a Namespace of the [main] fields) + the finalised parameter arena."""
    from neuralmonkey_b200 import runtime
    from neuralmonkey_b200.config.configuration import Configuration
    runtime.reset()
    torch.manual_seed(SEED)
    path = tempfile.NamedTemporaryFile("w", suffix=".ini", delete=False)
    path.write(ini_text.format(**fmt))
    path.close()
    _TMP.append(path.name)
    config = Configuration()
    for name in fields:
        config.add_argument(name)
    config.load_file(path.name)
    config.build_model()
    model = config.model
    for part in model.trainer.parameterizeds:
        part.ensure_declared()
    for extra in ("bs_decoder",):
        if hasattr(model, extra):
            getattr(model, extra).ensure_declared()
    runtime.arena().finalize(runtime.device())
    model.arena = runtime.arena()
    return model


def cleanup() -> None:
    for name in _TMP:
        try:
            os.unlink(name)
        except OSError:
            pass
    del _TMP[:]


def build_ende(vocab: int = 32000, cuda_graph: bool = True, beam_steps: int = 128) -> Namespace:
    return build(ENDE_INI, ("encoder", "attention", "decoder", "trainer", "runner", "bs_decoder"),
                 vocab=_wordlist(vocab), graph=cuda_graph, beam_steps=beam_steps)


def build_transformer(vocab: int = 32000, max_len: int = 64, cuda_graph: bool = True, dropout: bool = False,
                      tie: bool = True, beam_steps: int = 128) -> Namespace:
    keep = 0.9 if dropout else 1.0
    return build(TRANSFORMER_INI, ("input_sequence", "encoder", "decoder", "trainer", "runner", "bs_decoder"),
                 vocab=_wordlist(vocab), max_len=max_len, keep=keep, graph=cuda_graph, tie=tie,
                 beam_steps=beam_steps)


def build_captioning(vocab: int = 10000, max_len: int = 16, cuda_graph: bool = True) -> Namespace:
    return build(CAPTIONING_INI, ("encoder", "attention", "decoder", "trainer", "runner"),
                 vocab=_wordlist(vocab), max_len=max_len, graph=cuda_graph)


# ---- feeding id tensors (the stage behind pad_batch / strings_to_indices) ------------------------------
def feed_ende(model: Namespace, src_ids: torch.Tensor, tgt_ids, train: bool) -> None:
    """src_ids [B,Tx], tgt_ids [B,Ty] incl. </s> (or None): int64 host (pinned) or device tensors."""
    bsz = src_ids.shape[0]
    enc, att, dec = model.encoder, model.attention, model.decoder
    enc.input_sequence.feed_ids([src_ids], train=train)
    for part in (enc, att):
        part.reset_batch()
        part.train_mode = train
        part.batch_size = bsz
    dec.feed_ids(tgt_ids, bsz, train=train)
    if hasattr(model, "bs_decoder"):
        model.bs_decoder.reset_batch()
        model.bs_decoder.batch_size = bsz


def feed_transformer(model: Namespace, src_ids: torch.Tensor, tgt_ids, train: bool) -> None:
    bsz = src_ids.shape[0]
    model.input_sequence.feed_ids([src_ids], train=train)
    enc = model.encoder
    enc.reset_batch()
    enc.train_mode = train
    enc.batch_size = bsz
    model.decoder.feed_ids(tgt_ids, bsz, train=train)
    if hasattr(model, "bs_decoder"):
        model.bs_decoder.reset_batch()
        model.bs_decoder.batch_size = bsz


def feed_captioning(model: Namespace, images: torch.Tensor, tgt_ids, train: bool) -> None:
    bsz = images.shape[0]
    model.encoder.feed_images(images, train=train)
    att = model.attention
    att.reset_batch()
    att.train_mode, att.batch_size = train, bsz
    model.decoder.feed_ids(tgt_ids, bsz, train=train)


def synthetic_ids(bsz: int, length: int, vocab: int, seed: int, eos: bool = True) -> torch.Tensor:
    """Token ids uniform over [4, V) (never specials); targets end with </s>."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(4, vocab, (bsz, length), generator=g)
    if eos:
        ids[:, -1] = 2
    return ids
