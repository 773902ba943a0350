"""Host-side logic next to the hot path: dataset batching, learning-rate
schedule, BLEU, image reader, beam runner ranges - all on the CPU."""
import math
import os

import numpy as np
import pytest


def _write(path, lines):
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def test_dataset_fixed_batches_and_series(tmp_path):
    from neuralmonkey_b200.dataset import BatchingScheme, load
    src, tgt = tmp_path / "s.txt", tmp_path / "t.txt"
    _write(src, ["a b c", "d e", "f", "g h i j", "k"])
    _write(tgt, ["A B", "C", "D E F", "G", "H I"])
    ds = load("toy", ["source", "target"], [str(src), str(tgt)], BatchingScheme(batch_size=2))
    batches = list(ds.batches())
    assert [len(b) for b in batches] == [2, 2, 1]
    assert list(batches[0].get_series("source")) == [["a", "b", "c"], ["d", "e"]]
    assert list(batches[2].get_series("target")) == [["H", "I"]]
    dropped = load("toy", ["source", "target"], [str(src), str(tgt)],
                   BatchingScheme(batch_size=2, drop_remainder=True))
    assert [len(b) for b in dropped.batches()] == [2, 2]


def test_dataset_bucketing_groups_by_length(tmp_path):
    from neuralmonkey_b200.dataset import BatchingScheme, load
    src = tmp_path / "s.txt"
    _write(src, ["x"] * 5 + ["x y z w"] * 3 + ["x y z w u v t"] * 2)
    scheme = BatchingScheme(bucket_boundaries=[2, 5], bucket_batch_sizes=[4, 2, 1])
    ds = load("toy", ["source"], [str(src)], scheme)
    seen = 0
    for batch in ds.batches():
        lens = [len(s) for s in batch.get_series("source")]
        seen += len(lens)
        buckets = {0 if n <= 2 else (1 if n <= 5 else 2) for n in lens}
        assert len(buckets) == 1                          # never mixes buckets
        assert len(lens) <= [4, 2, 1][buckets.pop()]      # bucket batch size respected
    assert seen == 10


def test_batching_scheme_validation():
    from neuralmonkey_b200.dataset import BatchingScheme
    with pytest.raises(ValueError):
        BatchingScheme()
    with pytest.raises(ValueError):
        BatchingScheme(batch_size=4, bucket_boundaries=[3], bucket_batch_sizes=[1, 1])


def test_noam_decay_uses_step_before_increment():
    from neuralmonkey_b200.functions import noam_decay
    sched = noam_decay(learning_rate=0.2, model_dimension=6, warmup_steps=111)
    assert sched(0) == 0.0                                # first update runs at rate 0
    peak = 0.2 / math.sqrt(6) / math.sqrt(111)
    assert abs(sched(111) - peak) < 1e-12
    assert sched(50) < sched(111) > sched(500)
    assert abs(sched(444) - peak / 2) < 1e-12             # step^-0.5 decay after warm-up


def test_bleu_matches_hand_computation():
    from neuralmonkey_b200.evaluators.bleu import BLEUEvaluator
    bleu = BLEUEvaluator()
    ref = [["the", "cat", "sat", "on", "the", "mat"]]
    assert abs(bleu(ref, ref) - 100.0) < 1e-9
    # nothing matches: unigram precision 0 -> smoothed to 1 / (2 * 1); orders 2-4 have no hypothesis n-gram
    # and count as 1; brevity penalty exp(1 - 6/1)  (mteval-v13a smoothing, evaluators/bleu.py:196-208)
    import math
    assert bleu([["dog"]], ref) == pytest.approx(100 * math.exp(0.25 * math.log(0.5) + (1 - 6)))
    hyp = [["the", "cat", "sat", "on", "a", "mat"]]
    # precisions 5/6, 3/5, 1/4, 0/3 smoothed by the evaluator: just bounded here
    assert 0.0 <= bleu(hyp, ref) < 100.0
    assert BLEUEvaluator.compare_scores(2.0, 1.0) == 1


def test_imagenet_reader_crops_and_pads(tmp_path):
    """Reference quirk kept on purpose: no rescaling, centre crop + zero padding
    (readers/image_reader.py:147-178)."""
    from neuralmonkey_b200.readers.image_reader import imagenet_reader
    big = np.arange(10 * 12 * 3, dtype=np.float64).reshape(10, 12, 3)
    small = np.ones((4, 5, 3))
    np.save(tmp_path / "big.npy", big)
    np.save(tmp_path / "small.npy", small)
    _write(tmp_path / "list.txt", ["big.npy", "small.npy"])
    load = imagenet_reader(str(tmp_path), target_width=8, target_height=6, vgg_normalization=True)
    a, b = list(load([str(tmp_path / "list.txt")]))
    assert a.shape == b.shape == (6, 8, 3)
    means = np.array([123.68, 116.779, 103.939])
    assert np.allclose(a + means, big[2:8, 2:10])         # centre crop of the larger image
    assert np.allclose(b[:4, :5] + means, 1.0) and np.allclose(b[4:, :] + means, 0.0)


def test_beam_search_runner_range_names_and_validation():
    from neuralmonkey_b200.runners import beam_search_runner_range

    from neuralmonkey_b200.model.model_part import GenericModelPart

    class Dec(GenericModelPart):
        beam_size = 3
        vocabulary = None
    runners = beam_search_runner_range("target", Dec(), max_rank=2)
    assert [r.output_series for r in runners] == ["target.rank001", "target.rank002"]
    with pytest.raises(ValueError):
        beam_search_runner_range("target", Dec(), max_rank=4)


def test_bpe_matches_reference_subword_nmt(tmp_path):
    """Golden segmentations generated with the reference's vendored subword-nmt
    (tests/golden/make_bpe_golden.py)."""
    import json
    from neuralmonkey_b200.processors.bpe import BPEPostprocessor, BPEPreprocessor
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                         "bpe_golden.json")))
    merges = tmp_path / "merges.txt"
    _write(merges, golden["merges"])
    pre = BPEPreprocessor(str(merges))
    for sent, want in zip(golden["sentences"], golden["segmented"]):
        assert pre(sent) == want
    post = BPEPostprocessor()
    assert post([golden["segmented"][0]]) == [golden["sentences"][0]]


def test_plugin_names_of_the_reference_configs_are_importable():
    """Every `class=` path the five named INIs use resolves in this package."""
    import importlib
    names = """encoders.recurrent.SentenceEncoder encoders.SentenceEncoder encoders.transformer.TransformerEncoder
    encoders.imagenet_encoder.ImageNet model.sequence.EmbeddedSequence attention.Attention decoders.decoder.Decoder
    decoders.Decoder decoders.transformer.TransformerDecoder decoders.output_projection.maxout_output
    decoders.beam_search_decoder.BeamSearchDecoder trainers.cross_entropy_trainer.CrossEntropyTrainer
    trainers.cross_entropy_trainer.CostObjective trainers.delayed_update_trainer.DelayedUpdateTrainer
    trainers.multitask_trainer.MultitaskTrainer runners.GreedyRunner runners.beam_search_runner_range
    runners.BeamSearchRunner runners.runner.GreedyRunner runners.tensor_runner.TensorRunner
    runners.tensor_runner.RepresentationRunner tf_manager.TensorFlowManager dataset.load dataset.BatchingScheme
    vocabulary.from_wordlist functions.noam_decay evaluators.BLEU evaluators.ROUGE_L evaluators.SacreBLEU
    evaluators.bleu.BLEU1 evaluators.bleu.BLEU4 readers.image_reader.imagenet_reader
    processors.helpers.preprocess_char_based processors.bpe.BPEPreprocessor processors.bpe.BPEPostprocessor
    dataset.load_dataset_from_files vocabulary.from_bpe""".split()
    for name in names:
        module, attr = name.rsplit(".", 1)
        assert hasattr(importlib.import_module("neuralmonkey_b200." + module), attr), name
    from neuralmonkey_b200 import tf
    assert tf.contrib.opt.LazyAdamOptimizer and tf.train.AdamOptimizer


def test_tf_manager_keeps_n_best_checkpoints_and_restores(tmp_path):
    """n-best bookkeeping of TensorFlowManager (tf_manager.py:133-155 of the reference): the worst
    kept file is overwritten, `<prefix>.best` names the best one, restore brings the values back."""
    import torch
    from neuralmonkey_b200 import runtime
    from neuralmonkey_b200.params import normal_initializer
    from neuralmonkey_b200.tf_manager import TensorFlowManager
    runtime.reset()
    arena = runtime.arena()
    arena.declare("part/kernel", [4, 3], normal_initializer(0.5))
    arena.declare("part/bias", [3], normal_initializer(0.5))
    arena.finalize(torch.device("cpu"), seed=3)
    mgr = TensorFlowManager(num_sessions=1, num_threads=1, save_n_best=2)
    prefix = str(tmp_path / "variables.data")
    mgr.init_saving(prefix)
    assert mgr.variables_files == [prefix + ".0", prefix + ".1"]
    snapshots = {}
    for step, score in enumerate([0.1, 0.5, 0.3, 0.2, 0.9]):
        with torch.no_grad():
            arena.params.add_(1.0)
        snapshots[score] = arena.state_dict()
        mgr.validation_hook(score, epoch=1, batch=step)
    assert sorted(mgr.saved_scores) == [0.5, 0.9]
    assert mgr.best_score == 0.9
    best_name = open(prefix + ".best").read()
    assert best_name == os.path.basename(mgr.variables_files[mgr.best_score_index])
    with torch.no_grad():
        arena.params.zero_()
    mgr.restore_best_vars()
    for name, want in snapshots[0.9].items():
        assert torch.equal(arena.state_dict()[name], want)
    runtime.reset()


def test_rouge_l_and_accuracy_evaluators():
    from neuralmonkey_b200.evaluators import ROUGE_L, AccuracyEvaluator, AccuracySeqLevelEvaluator
    ref = [["a", "b", "c", "d"]]
    assert abs(ROUGE_L(ref, ref) - 1.0) < 1e-9
    assert ROUGE_L([["x"]], ref) == 0.0
    partial = ROUGE_L([["a", "c", "x"]], ref)
    assert 0.0 < partial < 1.0
    acc = AccuracyEvaluator()
    assert acc([["a", "b"], ["c"]], [["a", "x"], ["c"]]) == pytest.approx(2.0 / 3.0)     # token level
    assert AccuracySeqLevelEvaluator()([["a", "b"], ["c"]], [["a", "x"], ["c"]]) == pytest.approx(0.5)
    assert AccuracyEvaluator(mask_symbol="x")([["a", "b"]], [["a", "x"]]) == pytest.approx(1.0)


def test_sacrebleu_evaluator_known_answers():
    """evaluators.SacreBLEU follows sacrebleu's corpus BLEU (what the reference's wrapper calls with
    tokenize="none", smooth_method="exp"): a hand-computed corpus - precisions 7/8, 3/6, 2/4, 1/3, brevity
    penalty exp(1 - 9/8) - the exp smoothing of an order without matches, and the 13a tokenizer."""
    import math
    from neuralmonkey_b200.evaluators import SacreBLEU
    from neuralmonkey_b200.evaluators.sacrebleu import SacreBLEUEvaluator, tokenize_13a
    hyp = [["the", "cat", "sat", "on", "the", "mat"], ["hello", "world"]]
    ref = [["the", "cat", "sat", "on", "a", "mat"], ["hello", "there", "world"]]
    want = 100.0 * math.exp(1 - 9 / 8) * math.exp((math.log(7 / 8) + math.log(3 / 6) + math.log(2 / 4) + math.log(1 / 3)) / 4)
    assert abs(SacreBLEU(hyp, ref) - want) < 1e-9
    assert abs(SacreBLEU(ref, ref) - 100.0) < 1e-9
    assert SacreBLEU.name == "BLEU"
    # no 4-gram (and no trigram) match: exp smoothing gives 1/(2 t3) and 1/(4 t4) instead of zero
    short_h, short_r = [["a", "b", "c", "d"]], [["a", "b", "x", "d"]]
    want = 100.0 * math.exp((math.log(3 / 4) + math.log(1 / 3) + math.log(1 / (2 * 2)) + math.log(1 / (4 * 1))) / 4)
    assert abs(SacreBLEU(short_h, short_r) - want) < 1e-9
    assert SacreBLEUEvaluator("x", smooth_method="none")(short_h, short_r) < 1e-6
    assert tokenize_13a('Hello, world! It costs 3.5$ (approx.) - ok?') == "Hello , world ! It costs 3.5 $ ( approx . ) - ok ?"
    import pytest
    with pytest.raises(ValueError):
        SacreBLEUEvaluator("x", tokenize="intl-unknown")
