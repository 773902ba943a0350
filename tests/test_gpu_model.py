"""The Bahdanau encoder-decoder through the ModelPart API against the CPU oracle:
forward tensors, loss, every gradient, the optimizer step, and greedy decoding."""
import pytest
import torch

from oracle import nm_oracle as O
from tests.helpers import (build_bahdanau, feed, max_abs, oracle_params_for, oracle_spec,
                           random_batch, rel_err)

pytestmark = pytest.mark.gpu

TOY = dict(vs=60, vt=70, es=11, he=7, et=9, hd=8, out=9, maxout=True, max_len=10, supress_unk=True)
MID = dict(vs=120, vt=200, es=32, he=16, et=32, hd=32, out=32, maxout=False, max_len=12,
           supress_unk=False)


def _setup(cfg, backend, bsz=6, tx=8, ty=7, seed=0, **trainer_kw):
    from neuralmonkey_b200 import ops
    ops.set_gemm_backend(backend)
    model = build_bahdanau(**cfg, **trainer_kw)
    params = oracle_params_for(model)
    model["arena"].load_dict(params)
    src, tgt = random_batch(bsz, tx, ty, cfg["vs"], cfg["vt"], seed=seed)
    return model, params, src, tgt


@pytest.mark.parametrize("cfg,backend,tol", [(TOY, "simt", 2e-5), (MID, "simt", 2e-5),
                                             (MID, "auto", 1e-2)])
def test_train_forward_and_gradients(cfg, backend, tol):
    from neuralmonkey_b200 import ops
    try:
        model, params, src, tgt = _setup(cfg, backend)
        feed(model, src, tgt, train=True)
        enc, dec = model["enc"], model["dec"]
        spec = oracle_spec(cfg["maxout"], cfg["max_len"], cfg["supress_unk"])
        p64 = {n: v.double().requires_grad_(True) for n, v in params.items()}
        oenc = O.sentence_encoder(p64, "sentence_encoder", src)
        odec = O.decoder_train(p64, spec, oenc, tgt.t())
        assert max_abs(enc.temporal_states, oenc["temporal_states"]) < tol
        assert max_abs(enc.output, oenc["output"]) < tol
        assert max_abs(dec.train_output_states, odec["train_output_states"]) < tol
        assert max_abs(dec.train_xents, odec["train_xents"]) < 10 * tol
        assert abs(float(dec.train_loss) - float(odec["train_loss"])) < max(tol, 1e-5)
        keep = torch.ones(cfg["vt"], dtype=torch.bool)
        keep[3] = not cfg["supress_unk"]  # the suppressed <unk> column holds -1e9
        assert max_abs(dec.train_logits.cpu()[..., keep], odec["train_logits"][..., keep]) < 10 * tol
        # gradients of the token-mean loss w.r.t. every variable
        arena = model["arena"]
        arena.zero_grad()
        dec.train_loss.backward()
        odec["train_loss"].backward()
        gtol = 2e-4 if backend == "simt" else 1e-2
        for name, grad in arena.named_grads().items():
            want = p64[name].grad
            want = torch.zeros_like(p64[name]) if want is None else want
            err = float((grad.double() - want.reshape(grad.shape)).norm())
            assert err <= gtol * float(want.norm()) + 1e-7, (name, err, float(want.norm()))
    finally:
        ops.set_gemm_backend("auto")


@pytest.mark.parametrize("trainer_kw", [dict(), dict(l2=1e-3, clip=1.0), dict(l1=1e-4, clip=10.0)])
def test_optimizer_steps_match_oracle(trainer_kw):
    """Three CrossEntropyTrainer steps: losses and all parameters track TF-Adam semantics."""
    from neuralmonkey_b200 import ops
    try:
        model, params, src, tgt = _setup(TOY, "simt", lr=1e-2, **trainer_kw)
        spec = oracle_spec()
        p32 = {n: v.clone() for n, v in params.items()}
        st = O.AdamState(p32)
        for step in range(3):
            src, tgt = random_batch(6, 8, 7, TOY["vs"], TOY["vt"], seed=10 + step)
            feed(model, src, tgt, train=True)
            out = model["trainer"].train_step()
            ref = O.train_step(p32, spec, "sentence_encoder", src, tgt.t(), st,
                               l1=trainer_kw.get("l1", 0.0), l2=trainer_kw.get("l2", 0.0),
                               clip_norm=trainer_kw.get("clip"), lr=1e-2)
            assert abs(float(out["losses"][0]) - float(ref["loss"])) < 2e-4
            if trainer_kw:
                assert abs(float(out["l1l2"][1]) - float(ref["l2"])) < 1e-3 * float(ref["l2"]) + 1e-5
        got = model["arena"].state_dict()
        for name, want in p32.items():
            if name.endswith("attn_bias"):
                continue   # softmax is shift-invariant: its gradient is rounding noise, which Adam normalises to +-lr
            assert max_abs(got[name], want) < 5e-4, name
    finally:
        ops.set_gemm_backend("auto")


@pytest.mark.parametrize("cfg,backend,tol", [(TOY, "simt", 5e-5), (MID, "auto", 2e-2)])
def test_greedy_decoding(cfg, backend, tol):
    from neuralmonkey_b200 import ops
    try:
        model, params, src, tgt = _setup(cfg, backend, seed=3)
        feed(model, src, tgt, train=False)
        dec = model["dec"]
        spec = oracle_spec(cfg["maxout"], cfg["max_len"], cfg["supress_unk"])
        oenc = O.sentence_encoder(params, "sentence_encoder", src)
        og = O.decoder_greedy(params, spec, oenc, tgt.t())
        steps = og["runtime_logits"].shape[0]
        assert dec.runtime_logits.shape[0] == steps
        keep = torch.ones(cfg["vt"], dtype=torch.bool)
        keep[3] = not cfg["supress_unk"]
        assert max_abs(dec.runtime_logits.cpu()[..., keep], og["runtime_logits"][..., keep]) < tol
        if backend == "simt":   # integer bookkeeping is exact when the logits are fp32-exact
            assert bool((dec.runtime_symbols.cpu() == og["output_symbols"]).all())
            assert bool((dec.runtime_mask.cpu() == og["runtime_mask"]).all())
            assert bool((dec.decoded.cpu() == og["decoded"]).all())
        assert abs(float(dec.runtime_loss) - float(og["runtime_loss"])) < 20 * tol
        assert max_abs(dec.runtime_logprobs.cpu()[..., keep], og["runtime_logprobs"][..., keep]) < 2 * tol
    finally:
        ops.set_gemm_backend("auto")


def test_greedy_runner_tokens():
    from neuralmonkey_b200 import ops
    from neuralmonkey_b200.runners import GreedyRunner
    try:
        model, params, src, tgt = _setup(TOY, "simt", seed=4)
        feed(model, src, tgt, train=False)
        runner = GreedyRunner(output_series="target", decoder=model["dec"])
        exe = runner.get_executable(compute_losses=True, summaries=False, num_sessions=1)
        exe.execute()
        res = exe.result
        oenc = O.sentence_encoder(params, "sentence_encoder", src)
        og = O.decoder_greedy(params, oracle_spec(), oenc, tgt.t())
        want = model["dec"].vocabulary.vectors_to_sentences(
            og["runtime_logprobs"].argmax(-1).numpy())
        assert res.outputs["target"] == want
        assert abs(res.losses["target/runtime_xent"] - float(og["runtime_loss"])) < 1e-3
    finally:
        ops.set_gemm_backend("auto")


def test_cuda_graph_training_step_matches_eager():
    """use_cuda_graph=True: the first batch of a shape runs eagerly, the second is captured, later
    ones replay; losses and parameters follow the eager trainer step for step."""
    from neuralmonkey_b200 import ops
    try:
        results = {}
        for mode in (False, True):
            ops.set_gemm_backend("simt")
            model = build_bahdanau(**TOY, lr=1e-2, clip=1.0, l2=1e-3, cuda_graph=mode)
            params = oracle_params_for(model)
            model["arena"].load_dict(params)
            losses = []
            for step in range(5):
                src, tgt = random_batch(6, 8, 7, TOY["vs"], TOY["vt"], seed=30 + step, ragged=False)
                feed(model, src, tgt, train=True)
                losses.append(float(model["trainer"].train_step()["losses"][0]))
            results[mode] = (losses, model["arena"].state_dict())
            if mode:
                captured = [v for v in model["trainer"]._graphs.values() if isinstance(v, tuple)]
                assert len(captured) == 1, model["trainer"]._graphs
        print("eager", results[False][0], "graph", results[True][0])
        for a, b in zip(results[False][0], results[True][0]):
            assert abs(a - b) < 1e-5
        for name, want in results[False][1].items():
            if name.endswith("attn_bias"):
                continue   # zero-mean noise gradient (see test_optimizer_steps_match_oracle)
            assert max_abs(results[True][1][name], want) < 2e-5, name
    finally:
        ops.set_gemm_backend("auto")


def test_lazy_adam_leaves_absent_embedding_rows_untouched():
    """tf.contrib.opt.LazyAdamOptimizer: after a first step that touched every row, a second step
    on a batch using few words must not move the rows of the words it does not contain (dense Adam
    would keep moving them along their first moment)."""
    from neuralmonkey_b200 import ops, tf
    try:
        ops.set_gemm_backend("simt")
        results = {}
        for lazy in (True, False):
            model = build_bahdanau(**TOY, lr=1e-2)
            opt = tf.contrib.opt.LazyAdamOptimizer(learning_rate=1e-2) if lazy else tf.AdamOptimizer(learning_rate=1e-2)
            model["trainer"].optimizer = opt
            model["arena"].load_dict(oracle_params_for(model))
            bsz = (TOY["vs"] - 4) // 2
            src = torch.arange(4, TOY["vs"]).view(bsz, 2)                           # every source word
            tgt = torch.full((bsz, 3), 5)
            tgt[:, 2] = 2
            feed(model, src, tgt, train=True)
            model["trainer"].train_step()
            before = model["arena"].state_dict()["sentence_encoder_input/embedding_matrix_0"].clone()
            src2 = torch.full((4, 2), 7)
            feed(model, src2, tgt[:4], train=True)
            model["trainer"].train_step()
            after = model["arena"].state_dict()["sentence_encoder_input/embedding_matrix_0"]
            results[lazy] = (after - before).abs().sum(dim=1)
        moved_lazy, moved_dense = results[True], results[False]
        assert float(moved_lazy[7]) > 0                      # the word in the batch is updated
        assert float(moved_lazy[20:40].max()) == 0.0         # absent rows: untouched
        assert float(moved_dense[20:40].min()) > 0.0         # dense Adam keeps moving them
    finally:
        ops.set_gemm_backend("auto")
