"""Golden vectors from the reference's own runner post-processing (host side, numpy only), run here:

  runners/runner.py            GreedyRunner.Executable.collect_results   (1 and 2 sessions)
  runners/beamsearch_runner.py BeamSearchRunner.Executable.prepare_results, _is_finished
  runners/base_runner.py       set_runner_result (names of the losses)

    python tests/golden/make_runner_golden.py   ->  tests/golden/runner_golden.json
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim as shim  # noqa: E402
from make_host_golden import install_stubs  # noqa: E402

WORDS = ["the", "cat", "sat", "on", "mat", "a", "dog", "barks", "hello", "world"]


def main():
    install_stubs()
    shim.install()
    from neuralmonkey import vocabulary as V
    from neuralmonkey.runners.runner import GreedyRunner
    from neuralmonkey.runners.beamsearch_runner import BeamSearchRunner
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "vocab.tsv")
        with open(path, "w") as f:
            for w in ["<pad>", "<s>", "</s>", "<unk>"] + WORDS:
                f.write(w + "\n")
        vocab = V.from_wordlist(path, contains_header=False, contains_frequencies=False)
    vsz = len(vocab.index_to_word)
    rng = np.random.RandomState(5)
    out = {"words": WORDS}

    def logprobs(steps, bsz):
        x = rng.randn(steps, bsz, vsz).astype(np.float32) * 2.0
        x[rng.randint(steps), :, 2] += 6.0               # an early </s> somewhere
        return x - np.log(np.exp(x).sum(-1, keepdims=True))

    # ---- greedy runner ------------------------------------------------------------------------------
    def upper(sentences):
        return [[w.upper() for w in s] for s in sentences]

    for name, n_sess, post in (("single", 1, None), ("ensemble", 2, None), ("post", 1, upper)):
        results = [{"decoded_logprobs": logprobs(5, 3), "train_xent": float(rng.rand()),
                    "runtime_xent": float(rng.rand())} for _ in range(n_sess)]
        ex = object.__new__(GreedyRunner.Executable)
        ex._executor = types.SimpleNamespace(vocabulary=vocab, postprocess=post, output_series="target",
                                             loss_names=["train_xent", "runtime_xent"])
        ex.collect_results(results)
        out["greedy_" + name] = {
            "logprobs": [r["decoded_logprobs"].tolist() for r in results],
            "train_xent": [r["train_xent"] for r in results], "runtime_xent": [r["runtime_xent"] for r in results],
            "outputs": ex.result.outputs["target"], "losses": ex.result.losses, "size": ex.result.size}

    # ---- beam search runner -------------------------------------------------------------------------------
    steps, bsz, beam = 6, 3, 4
    token_ids = rng.randint(4, vsz, size=(steps + 1, bsz, beam)).astype(np.int64)
    token_ids[0] = 1
    token_ids[3, 0, :] = 2                   # sentence 0: every hypothesis ends after two words
    token_ids[1, 1, 1] = 2                   # sentence 1, rank 2: an EMPTY hypothesis
    token_ids[6, 2, 0] = 2
    scores = np.sort(rng.randn(bsz, beam).astype(np.float32), axis=1)[:, ::-1].copy()
    out["beam"] = {"token_ids": token_ids.tolist(), "scores": scores.tolist(), "ranks": {}}
    for rank in (1, 2, 4):
        ex = object.__new__(BeamSearchRunner.Executable)
        ex._executor = types.SimpleNamespace(output_series="target.rank{:03d}".format(rank), loss_names=["beam_search_score"])
        ex.rank, ex.postprocess = rank, None
        ex.decoder = types.SimpleNamespace(vocabulary=vocab)
        ex.prepare_results(types.SimpleNamespace(scores=scores, token_ids=token_ids))
        outputs = ex.result.outputs[ex._executor.output_series]
        out["beam"]["ranks"][str(rank)] = {
            # the reference leaves the raw id array in place of an empty hypothesis (see the product's docstring)
            "outputs": [o if isinstance(o, list) else {"raw_ids": np.asarray(o).tolist()} for o in outputs],
            "losses": {k: float(v) for k, v in ex.result.losses.items()}, "size": ex.result.size}

    # _is_finished: all sessions' decoders finished, or the step count reached max_steps
    cases = []
    for finished, n_tok, max_steps in (([True, True], 3, 10), ([True, False], 3, 10), ([False, False], 11, 10),
                                       ([False, False], 10, 10)):
        ex = object.__new__(BeamSearchRunner.Executable)
        ex.decoder = types.SimpleNamespace(max_steps_int=max_steps)
        res = [{"bs_outputs": types.SimpleNamespace(
            last_dec_loop_state=types.SimpleNamespace(feedables=types.SimpleNamespace(finished=np.array(finished))),
            last_search_step_output=types.SimpleNamespace(token_ids=np.zeros((n_tok, 1, 1))))}]
        cases.append({"finished": finished, "n_token_rows": n_tok, "max_steps": max_steps,
                      "is_finished": bool(ex._is_finished(res))})
    out["beam_is_finished"] = cases

    # ---- learning_utils.py: result joining, evaluation dictionary, log line formats, the validation preview ----
    from collections import OrderedDict
    from neuralmonkey import learning_utils as LU
    from neuralmonkey.runners.base_runner import ExecutionResult
    printed = []
    LU.log_print = printed.append
    LU.log = lambda message, color=None: printed.append(message)
    results = [ExecutionResult({"target": [["a", "b"], ["c"]]}, {"target/xent": 2.0}, 2, []),
               ExecutionResult({"target": [["d"]]}, {"target/xent": 5.0}, 1, [])]
    joined = LU.join_execution_results(results)
    out["lu_join"] = {"outputs": joined.outputs, "losses": joined.losses, "size": joined.size}
    arrays = [ExecutionResult({"enc": [np.ones(3), np.zeros(3)]}, {}, 2, []), ExecutionResult({"enc": [np.ones(3)]}, {}, 1, [])]
    out["lu_join_arrays_shape"] = list(LU.join_execution_results(arrays).outputs["enc"].shape)

    class Exact:
        name = "exact"

        def __call__(self, hyp, ref):
            return float(np.mean([h == r for h, r in zip(hyp, ref)]))
    batch = {"target": [["a", "b"], ["x"], ["d"]], "source": [["s1"], ["s2"], ["s3"]]}
    evaluated = LU.evaluation([("target", "target", Exact()), ("missing", "target", Exact()), ("target", "nothere", Exact())],
                              batch, [joined], {"target": joined.outputs["target"]})
    out["lu_evaluation"] = list(evaluated.items())
    line = OrderedDict([("target/xent", 3.0), ("target/BLEU", 12.3456789), ("target/exact", 2.0 / 3), ("big", 123456.789)])
    out["lu_format_line"] = LU._format_evaluation_line(line, "target/BLEU")
    del printed[:]
    LU.print_final_evaluation(line, "test_0")
    out["lu_final_evaluation"] = list(printed)
    items = [["a", "b"], {"k": ["v", "w"], "n": 3}, np.zeros((2, 3)), np.arange(3), 4.5, "text", [["x"], ["y", "z"]]]
    out["lu_data_item_to_str"] = [LU._data_item_to_str(i) for i in items]
    del printed[:]
    LU._print_examples({"source": [["s1"], ["s2", "s2"], ["s3"]], "target": [["t1"], ["t2"], ["t3"]], "extra": [1, 2, 3]},
                       {"target": [["o1"], ["o2"], ["o3"]], "rep": [np.zeros((2, 2)), np.zeros((2, 2)), np.zeros((2, 2))]},
                       num_examples=2)
    out["lu_examples_all"] = list(printed)
    del printed[:]
    LU._print_examples({"source": [["s1"]], "target": [["t1"]], "extra": [1]}, {"target": [["o1"]], "rep": [7]},
                       val_preview_input_series=["source", "target"], val_preview_output_series=["target"])
    out["lu_examples_selected"] = list(printed)
    with open(os.path.join(HERE, "runner_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["beam"]["ranks"], indent=1)[:1500])


if __name__ == "__main__":
    main()
