"""Golden vectors for the HOST side of the path, produced by the reference's own code run in this
container: vocabulary / padding / index conversion (`neuralmonkey/vocabulary.py`), dataset batching and
bucketing (`neuralmonkey/dataset.py`), BLEU (`neuralmonkey/evaluators/bleu.py`), char-level helpers
(`neuralmonkey/processors/helpers.py`).  Only third-party imports that are absent here are stubbed
(termcolor, typeguard, and an empty `tensorflow` module - none of the functions called below use it);
`collections.Sized/Iterable` get their Python >= 3.10 aliases.
    python tests/golden/make_host_golden.py
"""
import json
import os
import sys
import tempfile
import types

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

SENTENCES = [["the", "cat", "sat"], ["a", "dog"], [], ["the", "zebra", "sat", "on", "the", "mat", "again"]]
WORDS = ["the", "cat", "sat", "a", "dog", "on", "mat"]
HYPS = [["the", "cat", "sat", "on", "the", "mat"], ["a", "dog", "dog", "barks"], ["hello"], []]
REFS = [["the", "cat", "sat", "on", "a", "mat"], ["a", "dog", "barks"], ["hello", "world"], ["x"]]
CORPUS = ["a", "b c", "d e f", "g h i j", "k l m n o", "p", "q r", "s t u v w x y", "z z", "a b c d e f g h i"]


TEXT_LINES = ["Hello, world!  It's 42.", "  leading and trailing  ", "naive café: 3.14 -- ok", "x", "a+b=c (d)",
              "Ünïcödé ... \u4e2d\u6587 test"]
TSV_LINES = ["id1\tthe first text\tlabel a", "id2\t second  text \tlabel b", "id3\tthird"]
T2T_VOCAB_LINES = ["'<pad>'", "'<EOS>'", "'hello_'", '"quoted"', "plain", "'", "''", "'a\"", "  'spaced'  ", "<pad>",
                   "x'y", "'it''s'"]
XENT_ROWS = [[1.0, 3.0, 0.0], [2.0, 0.0, 0.0], [0.25, 0.5, 4.0]]
CSV_LINES = ['one two, "quoted, with comma", x y', 'three, plain field, z', 'four five,,']


def install_stubs():
    term = types.ModuleType("termcolor")
    term.colored = lambda text, *a, **k: text
    sys.modules["termcolor"] = term
    guard = types.ModuleType("typeguard")
    guard.check_argument_types = lambda *a, **k: True

    def matches(value, expected):
        import collections.abc
        import typing
        origin, args = typing.get_origin(expected), typing.get_args(expected)
        if expected is typing.Any:
            return True
        if origin is typing.Union:
            return any(matches(value, a) for a in args)
        if origin in (list, typing.List):
            return isinstance(value, list) and all(matches(v, args[0]) for v in value) if args else isinstance(value, list)
        if origin in (tuple, typing.Tuple):
            return isinstance(value, tuple) and len(value) == len(args) and all(matches(v, a) for v, a in zip(value, args))
        if origin in (dict, typing.Dict):
            return isinstance(value, dict) and all(matches(k, args[0]) and matches(v, args[1]) for k, v in value.items())
        if origin is collections.abc.Callable or expected is typing.Callable:
            return callable(value)
        return isinstance(value, expected)

    def check_type(_name, value, expected, _memo=None):
        if not matches(value, expected):
            raise TypeError("type mismatch")
    guard.check_type = check_type
    sys.modules["typeguard"] = guard
    tf = types.ModuleType("tensorflow")
    tf.Tensor = object
    # Vocabulary.__init__ builds two TF lookup tables we never query here
    lookup = types.SimpleNamespace(index_table_from_tensor=lambda *a, **k: None,
                                   index_to_string_table_from_tensor=lambda *a, **k: None)
    tf.contrib = types.SimpleNamespace(lookup=lookup)
    sys.modules["tensorflow"] = tf
    # packages the evaluators package imports at module level but BLEU does not use
    for name, attrs in (("sacrebleu", {"corpus_bleu": None, "TOKENIZERS": {"none": None, "13a": None, "intl": None, "zh": None}}), ("rouge", {"Rouge": object}),
                        ("pyter", {"ter": None})):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules[name] = mod
    # the reference predates Python 3.10: collections.Sized / Iterable moved to collections.abc
    import collections
    import collections.abc
    for name in ("Sized", "Iterable", "Callable"):
        if not hasattr(collections, name):
            setattr(collections, name, getattr(collections.abc, name))
    sys.path.insert(0, REF)


def main():
    install_stubs()
    import numpy as np
    out = {}
    # ---- vocabulary ------------------------------------------------------------------------
    from neuralmonkey import vocabulary as V
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "vocab.tsv")
        with open(path, "w") as f:
            for w in ["<pad>", "<s>", "</s>", "<unk>"] + WORDS:
                f.write(w + "\n")
        vocab = V.from_wordlist(path, contains_header=False, contains_frequencies=False)
    out["vocab_index_to_word"] = list(vocab.index_to_word)
    for max_len, start, end in ((None, False, False), (4, False, True), (3, True, True), (None, True, False)):
        padded = V.pad_batch([list(s) for s in SENTENCES], max_len, start, end)
        key = "pad_{}_{}_{}".format(max_len, int(start), int(end))
        out[key] = [list(s) for s in padded]
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "vocab.t2t")
        with open(path, "w") as f:
            f.write("\n".join(T2T_VOCAB_LINES) + "\n")
        out["t2t_vocabulary"] = list(V.from_t2t_vocabulary(path).index_to_word)
    out["sentence_mask_rule"] = "id != 0"
    vectors = np.array([[4, 5, 6, 2, 0], [7, 8, 2, 0, 0], [3, 3, 3, 3, 3]]).T      # time-major
    out["vectors_to_sentences"] = vocab.vectors_to_sentences(vectors)
    # ---- dataset batching --------------------------------------------------------------------
    from neuralmonkey.dataset import BatchingScheme, load
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "src.txt")
        with open(path, "w") as f:
            f.write("\n".join(CORPUS) + "\n")
        for name, scheme in (("fixed3", BatchingScheme(batch_size=3)),
                             ("fixed4_drop", BatchingScheme(batch_size=4, drop_remainder=True)),
                             ("buckets", BatchingScheme(bucket_boundaries=[2, 5],
                                                        bucket_batch_sizes=[3, 2, 1]))):
            ds = load("toy", ["source"], [path], scheme)
            out["dataset_" + name] = [[list(s) for s in b.get_series("source")] for b in ds.batches()]
        # series-level preprocessor, lazy buffer and shuffling (Python's `random`, seeded here)
        import random
        from neuralmonkey.processors.helpers import preprocess_char_based
        ds = load("toy", ["source", "chars"], [path, (preprocess_char_based, "source")], BatchingScheme(batch_size=4))
        out["dataset_preprocessed"] = [{k: [list(s) for s in b.get_series(k)] for k in ("source", "chars")}
                                       for b in ds.batches()]
        for name, kwargs in (("lazy", dict(buffer_size=4)), ("shuffled", dict(shuffled=True)),
                             ("lazy_shuffled", dict(buffer_size=6, shuffled=True))):
            random.seed(5)
            ds = load("toy", ["source"], [path], BatchingScheme(batch_size=3), **kwargs)
            out["dataset_" + name] = [[list(s) for s in b.get_series("source")] for b in ds.batches()]
    # ---- BLEU ------------------------------------------------------------------------------------
    from neuralmonkey.evaluators.bleu import BLEUEvaluator
    for n in (1, 2, 4):
        for dedup in (False, True):
            ev = BLEUEvaluator(n=n, deduplicate=dedup)
            out["bleu_{}_{}".format(n, int(dedup))] = float(ev(HYPS, REFS))
    out["bleu_identity"] = float(BLEUEvaluator()(REFS, REFS))
    # the corpus of the reference's own unit test (neuralmonkey/tests/test_bleu.py): the 4-gram precision is
    # zero, so the mteval-v13a smoothing acts; plus empty sides and several references per sentence
    ut_hyp = [d.split() for d in ("colorful thoughts furiously sleep", "little piglet slept all night",
                                  "working working working working working be be be be be be be", "ich bin walrus",
                                  "walrus for pr\u00e4sident")]
    ut_ref = [r.split() for r in ("the colorless ideas slept furiously", "pooh slept all night",
                                  "working class hero is something to be", "I am the working class walrus",
                                  "walrus for president")]
    multi_ref = [["a", "b", "|", "a", "c", "d"], ["x", "|", "x", "y"]]
    multi_hyp = [["a", "c"], ["x", "x", "y"]]
    out["bleu_unit_test"] = {
        "hyp": ut_hyp, "ref": ut_ref,
        "scores": {"{}_{}".format(n, int(d)): float(BLEUEvaluator(n=n, deduplicate=d)(ut_hyp, ut_ref))
                   for n in (1, 2, 4) for d in (False, True)},
        "empty_sentence": float(BLEUEvaluator()(ut_hyp + [["something"]], ut_ref + [[]])),
        "empty_decoded": float(BLEUEvaluator()([[] for _ in ut_hyp], ut_ref)),
        "empty_reference": float(BLEUEvaluator()(ut_hyp, [[] for _ in ut_ref])),
        "multi_hyp": multi_hyp, "multi_ref": multi_ref,
        "multi": float(BLEUEvaluator(n=2, multiple_references_separator="|")(multi_hyp, multi_ref))}
    # ---- further evaluators (chrf.py, edit_distance.py, mse.py, average.py; wer.py and ter.py need the
    #      third-party pyter and are not run) ------------------------------------------------------------
    from neuralmonkey.evaluators.chrf import ChrF3, ChrFEvaluator
    from neuralmonkey.evaluators.edit_distance import EditDistance
    from neuralmonkey.evaluators.mse import MSE, PairwiseMSE
    from neuralmonkey.evaluators.average import AverageEvaluator
    out["more_evaluators"] = {
        "chrf3": float(ChrF3(HYPS, REFS)), "chrf3_name": ChrF3.name,
        "chrf_default": float(ChrFEvaluator()(HYPS, REFS)),
        "chrf_ignored": float(ChrFEvaluator(n=3, beta=2.0, ignored_symbols=[" ", "a"])(HYPS, REFS)),
        "chrf_per_sentence": [float(ChrF3.score_instance(h, r)) for h, r in zip(HYPS + [[]], REFS + [[]])],
        "edit_distance": float(EditDistance(HYPS, REFS)), "edit_distance_name": EditDistance.name,
        "mse": float(MSE([[1.0, 2.0, 3.0], [0.5, 0.5, 0.5]], [[1.5, 2.0, 1.0], [0.0, 1.0, 0.5]])), "mse_name": MSE.name,
        "pairwise_mse": float(PairwiseMSE([[1.0, 2.0, 3.0], [0.5]], [[1.5, 2.0, 1.0], [0.0]])),
        "pairwise_mse_name": PairwiseMSE.name,
        "average": float(AverageEvaluator("avg")([1.0, 2.5, 4.0], [0.0, 0.0, 0.0]))}
    from neuralmonkey.evaluators.perplexity import PerplexityEvaluator
    out["more_evaluators"]["perplexity"] = float(PerplexityEvaluator("perplexity")(XENT_ROWS, [[], [], []]))
    out["more_evaluators"]["perplexity_name"] = PerplexityEvaluator("perplexity").name
    # ---- helpers -----------------------------------------------------------------------------------
    from neuralmonkey.processors import helpers as H
    out["char_based"] = [H.preprocess_char_based(s) for s in SENTENCES]
    out["char_based_back"] = H.postprocess_char_based(out["char_based"])
    # ---- text readers ------------------------------------------------------------------------------
    from neuralmonkey.readers import plain_text_reader as R
    with tempfile.TemporaryDirectory() as tmp:
        txt = os.path.join(tmp, "text.txt")
        with open(txt, "w", encoding="utf-8") as f:
            f.write("\n".join(TEXT_LINES) + "\n")
        tsv = os.path.join(tmp, "table.tsv")
        with open(tsv, "w", encoding="utf-8") as f:
            f.write("\n".join(TSV_LINES) + "\n")
        csvf = os.path.join(tmp, "table.csv")
        with open(csvf, "w", encoding="utf-8") as f:
            f.write("\n".join(CSV_LINES) + "\n")
        out["reader_tokenized"] = [list(x) for x in R.tokenized_text_reader()([txt])]
        out["reader_t2t"] = [list(x) for x in R.t2t_tokenized_text_reader()([txt])]
        out["reader_tsv2"] = [list(x) for x in R.tsv_reader(2)([tsv])]
        out["reader_csv1"] = [list(x) for x in R.csv_reader(1)([csvf])]
        out["reader_csv3"] = [list(x) for x in R.csv_reader(3)([csvf])]
    # ---- writers ---------------------------------------------------------------------------------------
    from neuralmonkey.writers import plain_text_writer as W
    from neuralmonkey.writers.auto import AutoWriter
    with tempfile.TemporaryDirectory() as tmp:
        def text_of(writer, data, name):
            target = os.path.join(tmp, name)
            writer(target, data)
            return open(target, encoding="utf-8").read()
        out["writer_tokenized"] = text_of(W.tokenized_text_writer(), out["reader_tokenized"], "a.txt")
        out["writer_t2t"] = text_of(W.t2t_tokenized_text_writer(), out["reader_t2t"], "b.txt")
        out["writer_plain"] = text_of(W.text_writer(), ["x y", 3, 4.5], "c.txt")
        out["writer_auto_tokens"] = text_of(AutoWriter, [["a", "b"], ["c"]], "d.txt")
        out["writer_auto_plain"] = text_of(AutoWriter, [1.5, 2.5], "e.txt")
        AutoWriter(os.path.join(tmp, "f"), [{"x": np.ones((2, 3)), "y": np.zeros(4)}, {"x": np.ones((2, 3)), "y": np.ones(4)}])
        loaded = np.load(os.path.join(tmp, "f.npz"))
        out["writer_auto_npz"] = {k: list(loaded[k].shape) for k in loaded.files}
        AutoWriter(os.path.join(tmp, "g"), np.arange(6).reshape(2, 3))
        out["writer_auto_npy"] = np.load(os.path.join(tmp, "g.npy")).tolist()
    out["inputs"] = {"text_lines": TEXT_LINES, "tsv_lines": TSV_LINES, "csv_lines": CSV_LINES,
                     "sentences": SENTENCES, "words": WORDS, "hyps": HYPS, "refs": REFS, "corpus": CORPUS,
                     "t2t_vocab_lines": T2T_VOCAB_LINES, "xent_rows": XENT_ROWS}
    json.dump(out, open(os.path.join(HERE, "host_golden.json"), "w"), indent=1, sort_keys=True)
    print({k: (v if not isinstance(v, list) else "...") for k, v in out.items()})


if __name__ == "__main__":
    main()
