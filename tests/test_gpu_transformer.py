"""Transformer encoder / decoder and the beam search decoder through the ModelPart API
against the CPU oracle."""
import pytest
import torch

from oracle import nm_oracle as O
from tests.helpers import (build_bahdanau, feed, max_abs, oracle_params_for, oracle_spec,
                           random_batch)

pytestmark = pytest.mark.gpu

CFG = dict(vs=80, vt=96, dim=32, ff=64, depth=2, heads=4, max_len=9)


def build_transformer(vs, vt, dim, ff, depth, heads, max_len, tie=True, supress_unk=False):
    from neuralmonkey_b200 import runtime, tf
    from neuralmonkey_b200.decoders import TransformerDecoder
    from neuralmonkey_b200.encoders import TransformerEncoder
    from neuralmonkey_b200.model.sequence import EmbeddedSequence
    from neuralmonkey_b200.trainers import CrossEntropyTrainer
    from neuralmonkey_b200.vocabulary import Vocabulary

    runtime.reset()
    src_vocab = Vocabulary(["s{}".format(i) for i in range(vs - 4)])
    tgt_vocab = Vocabulary(["t{}".format(i) for i in range(vt - 4)])
    seq = EmbeddedSequence(name="input_sequence", vocabulary=src_vocab, data_id="source",
                           embedding_size=dim, max_length=max_len, scale_embeddings_by_depth=True)
    enc = TransformerEncoder(name="encoder", input_sequence=seq, ff_hidden_size=ff, depth=depth,
                             n_heads=heads)
    dec = TransformerDecoder(name="decoder", encoders=[enc], vocabulary=tgt_vocab, data_id="target",
                             ff_hidden_size=ff, n_heads_self=heads, n_heads_enc=heads, depth=depth,
                             max_output_len=max_len, embedding_size=dim, tie_embeddings=tie,
                             supress_unk=supress_unk)
    trainer = CrossEntropyTrainer(decoders=[dec], optimizer=tf.AdamOptimizer(learning_rate=1e-3))
    for part in trainer.parameterizeds:
        part.ensure_declared()
    runtime.arena().finalize(runtime.device())
    return {"seq": seq, "enc": enc, "dec": dec, "trainer": trainer, "arena": runtime.arena()}


def feed_transformer(model, src, tgt, train):
    bsz = src.shape[0]
    model["seq"].feed_ids([src], train=train)
    enc = model["enc"]
    enc.reset_batch()
    enc.train_mode = train
    enc.batch_size = bsz
    model["dec"].feed_ids(tgt, bsz, train=train)


def oracle_encoder(p, src, cfg):
    emb = p["input_sequence/embedding_matrix_0"]
    mask = (src != 0).to(emb.dtype)
    inputs = emb[src] * (mask * (cfg["dim"] ** 0.5)).unsqueeze(-1)
    return O.transformer_encoder(p, "encoder", inputs, mask, cfg["depth"], cfg["heads"])


def _setup(backend, tie=True, supress_unk=False, bsz=5, tx=8, ty=7, seed=0):
    from neuralmonkey_b200 import ops
    ops.set_gemm_backend(backend)
    model = build_transformer(**CFG, tie=tie, supress_unk=supress_unk)
    params = oracle_params_for(model, scale=0.2)
    for name in params:  # LayerNorm scales near one, like a trained model
        if name.endswith("gamma"):
            params[name] = 1.0 + params[name]
    model["arena"].load_dict(params)
    src, tgt = random_batch(bsz, tx, ty, CFG["vs"], CFG["vt"], seed=seed)
    return model, params, src, tgt


def test_position_signal_matches_oracle():
    from neuralmonkey_b200.encoders.transformer import position_signal
    for dim, length in ((32, 9), (33, 4), (512, 50)):
        assert max_abs(position_signal(dim, length), O.position_signal(dim, length)) < 1e-6


@pytest.mark.parametrize("backend,tie,tol", [("simt", True, 5e-5), ("simt", False, 5e-5),
                                             ("auto", True, 3e-2)])
def test_train_forward_and_gradients(backend, tie, tol):
    from neuralmonkey_b200 import ops
    try:
        model, params, src, tgt = _setup(backend, tie=tie, supress_unk=not tie)
        feed_transformer(model, src, tgt, train=True)
        enc, dec = model["enc"], model["dec"]
        spec = O.TransformerDecoderSpec("decoder", CFG["depth"], CFG["heads"], CFG["heads"],
                                        CFG["max_len"], tie, not tie)
        p64 = {n: v.double().requires_grad_(True) for n, v in params.items()}
        oenc = oracle_encoder(p64, src, CFG)
        odec = O.transformer_decoder_train(p64, spec, oenc, tgt)
        assert max_abs(enc.temporal_states, oenc["states"]) < tol
        assert max_abs(enc.output, oenc["output"]) < 10 * tol
        assert max_abs(dec.train_output_states.transpose(0, 1), odec["states"]) < tol
        assert max_abs(dec.train_xents, odec["xents"]) < 20 * tol
        assert abs(float(dec.train_loss) - float(odec["loss"])) < 10 * tol
        arena = model["arena"]
        arena.zero_grad()
        dec.train_loss.backward()
        odec["loss"].backward()
        gtol = 3e-4 if backend == "simt" else 2e-2
        for name, grad in arena.named_grads().items():
            want = p64[name].grad
            want = torch.zeros_like(p64[name]) if want is None else want
            err = float((grad.double() - want.reshape(grad.shape)).norm())
            assert err <= gtol * float(want.norm()) + 1e-6, (name, err, float(want.norm()))
    finally:
        ops.set_gemm_backend("auto")


def test_trainer_step_decreases_loss():
    from neuralmonkey_b200 import ops
    try:
        model, _params, src, tgt = _setup("auto")
        losses = []
        for _ in range(8):
            feed_transformer(model, src, tgt, train=True)
            losses.append(float(model["trainer"].train_step()["losses"][0]))
        assert losses[-1] < losses[0]
    finally:
        ops.set_gemm_backend("auto")


def test_greedy_decoding():
    from neuralmonkey_b200 import ops
    try:
        model, params, src, tgt = _setup("simt", seed=2)
        feed_transformer(model, src, tgt, train=False)
        dec = model["dec"]
        spec = O.TransformerDecoderSpec("decoder", CFG["depth"], CFG["heads"], CFG["heads"],
                                        CFG["max_len"], True, False)
        og = O.transformer_decoder_greedy(params, spec, oracle_encoder(params, src, CFG))
        assert dec.runtime_logits.shape == og["logits"].shape
        assert max_abs(dec.runtime_logits, og["logits"]) < 2e-4
        assert bool((dec.runtime_symbols.cpu() == og["symbols"]).all())
        assert bool((dec.runtime_mask.cpu() == og["mask"]).all())
    finally:
        ops.set_gemm_backend("auto")


def _oracle_transformer_beam(params, spec, oenc, beam, max_steps, alpha):
    emb = params["decoder/word_embeddings"]
    states = oenc["states"].repeat_interleave(beam, 0)
    emask = oenc["mask"].repeat_interleave(beam, 0)
    rows = states.shape[0]

    def run(seq, mask):
        out = O.transformer_decoder_stack(params, spec, seq, mask, states, emask)
        return torch.log_softmax(O.transformer_logits(params, spec, out[:, -1]), -1)

    seq0 = emb[torch.full((rows,), O.START, dtype=torch.int64)].unsqueeze(1)
    mask0 = torch.ones(rows, 1)
    first = run(seq0, mask0)

    def step_fn(state, words, finished):
        seq = torch.cat([state[0], emb[words].unsqueeze(1)], 1)
        mask = torch.cat([state[1], (~finished).to(emb.dtype).unsqueeze(1)], 1)
        return (seq, mask), run(seq, mask)

    return O.beam_search(step_fn, (seq0, mask0), first, beam, max_steps, alpha,
                         lambda st, idx: (st[0][idx], st[1][idx]))


@pytest.mark.parametrize("graph", [True, False])
@pytest.mark.parametrize("beam,alpha,bsz", [(1, 0.0, 3), (4, 1.0, 3), (5, 0.6, 1)])
def test_beam_search_transformer_parent(beam, alpha, bsz, graph):
    """graph=True: one CUDA graph replayed per step over static-shape buffers
    (decoders/beam_graph.py); graph=False: the step-by-step host loop.  Both against the oracle."""
    from neuralmonkey_b200 import ops
    from neuralmonkey_b200.decoders import BeamSearchDecoder
    from neuralmonkey_b200.runners import BeamSearchRunner
    try:
        model, params, src, tgt = _setup("simt", bsz=bsz, seed=5)
        bs = BeamSearchDecoder(name="bs", parent_decoder=model["dec"], beam_size=beam, max_steps=7,
                               length_normalization=alpha)
        bs.use_cuda_graph = graph
        bs.GRAPH_AFTER = 1        # capture on first use (default: second occurrence of a shape)
        feed_transformer(model, src, None, train=False)
        bs.reset_batch()
        bs.batch_size = bsz
        out = bs.outputs
        spec = O.TransformerDecoderSpec("decoder", CFG["depth"], CFG["heads"], CFG["heads"],
                                        CFG["max_len"], True, False)
        want = _oracle_transformer_beam(params, spec, oracle_encoder(params, src, CFG), beam, 7, alpha)
        got_tokens = out.last_search_step_output.token_ids.cpu()
        assert got_tokens.shape[0] == want["token_ids"].shape[0] + 1
        assert bool((got_tokens[1:] == want["token_ids"]).all())
        assert max_abs(out.last_search_step_output.scores, want["scores"]) < 2e-4
        assert bool((out.last_search_state.lengths.cpu() == want["lengths"]).all())
        assert bool((out.last_search_state.finished.cpu() == want["finished"]).all())
        # the runner cuts at </s> and maps to words
        runner = BeamSearchRunner(output_series="target", decoder=bs, rank=1)
        exe = runner.get_executable(compute_losses=False, summaries=False, num_sessions=1)
        exe.execute()
        sentences = exe.result.outputs["target"]
        vocab = model["dec"].vocabulary
        for b in range(bsz):
            toks = []
            for t in want["token_ids"][:, b, 0].tolist():
                if t == O.END:
                    break
                toks.append(vocab.index_to_word[t])
            assert sentences[b] == toks
    finally:
        ops.set_gemm_backend("auto")


@pytest.mark.parametrize("beam,bsz", [(3, 1), (4, 3)])
def test_beam_search_rnn_parent(beam, bsz):
    """Bahdanau decoder under beam search.  The reference supports batch size 1 only here
    (its attention broadcasts instead of tiling); batch 3 checks the tiled extension."""
    from neuralmonkey_b200 import ops
    from neuralmonkey_b200.decoders import BeamSearchDecoder
    try:
        ops.set_gemm_backend("simt")
        cfg = dict(vs=60, vt=70, es=11, he=7, et=9, hd=8, out=9, maxout=True, max_len=10,
                   supress_unk=True)
        model = build_bahdanau(**cfg)
        params = oracle_params_for(model)
        model["arena"].load_dict(params)
        src, _tgt = random_batch(bsz, 8, 7, cfg["vs"], cfg["vt"], seed=9)
        bs = BeamSearchDecoder(name="bs", parent_decoder=model["dec"], beam_size=beam, max_steps=8,
                               length_normalization=1.0)
        feed(model, src, None, train=False)
        bs.reset_batch()
        bs.batch_size = bsz
        out = bs.outputs

        spec = oracle_spec()
        oenc = O.sentence_encoder(params, "sentence_encoder", src)
        states = oenc["temporal_states"].repeat_interleave(beam, 0)
        mask = oenc["temporal_mask"].repeat_interleave(beam, 0)
        hidden = O.bahdanau_precompute(params, "attention", states)
        emb = params["decoder/word_embeddings"]
        prev0 = O.decoder_initial_state(params, spec, oenc["output"]).repeat_interleave(beam, 0)

        def run(embedded, prev):
            output, cell, _c, _w = O.decoder_step(params, spec, embedded, prev, hidden, states, mask)
            return cell, torch.log_softmax(O.state_to_logits(params, spec, output), -1)

        prev1, first = run(emb[torch.full((bsz * beam,), O.START, dtype=torch.int64)], prev0)

        def step_fn(prev, words, _finished):
            return run(emb[words], prev)

        want = O.beam_search(step_fn, prev1, first, beam, 8, 1.0, lambda st, idx: st[idx])
        got = out.last_search_step_output
        assert bool((got.token_ids.cpu()[1:] == want["token_ids"]).all())
        assert max_abs(got.scores, want["scores"]) < 1e-4
    finally:
        ops.set_gemm_backend("auto")


def test_kv_cache_equals_prefix_recompute():
    """The cached runtime step and the reference's recompute-the-prefix schedule give the same
    greedy symbols and (to fp32 rounding) the same logits."""
    from neuralmonkey_b200 import ops
    try:
        model, _params, src, tgt = _setup("simt", seed=7)
        dec = model["dec"]
        feed_transformer(model, src, tgt, train=False)
        cached_logits = dec.runtime_logits.clone()
        cached_symbols = dec.runtime_symbols.clone()
        dec.use_kv_cache = False
        feed_transformer(model, src, tgt, train=False)
        assert dec.runtime_loop_result.feedables.other.kv_cache is None
        assert bool((dec.runtime_symbols == cached_symbols).all())
        assert max_abs(dec.runtime_logits, cached_logits) < 1e-4
    finally:
        ops.set_gemm_backend("auto")


def test_beam_graph_reused_across_batches_and_early_stop():
    """The captured graph is replayed for a second batch of the same shape (fresh encoder states,
    same buffers), and a search whose hypotheses all finish early is trimmed to the step the
    reference loop stops at."""
    from neuralmonkey_b200 import ops
    from neuralmonkey_b200.decoders import BeamSearchDecoder
    try:
        model, params, src, _tgt = _setup("simt", bsz=2, seed=11)
        # make </s> very likely so the search ends long before max_steps
        params = dict(params)
        params["decoder/word_embeddings"] = params["decoder/word_embeddings"].clone()
        model["arena"].load_dict(params)
        spec = O.TransformerDecoderSpec("decoder", CFG["depth"], CFG["heads"], CFG["heads"],
                                        CFG["max_len"], True, False)
        bs_graph = BeamSearchDecoder(name="bs", parent_decoder=model["dec"], beam_size=3, max_steps=30,
                                     length_normalization=1.0)
        bs_loop = BeamSearchDecoder(name="bs2", parent_decoder=model["dec"], beam_size=3, max_steps=30,
                                    length_normalization=1.0)
        bs_loop.use_cuda_graph = False
        for seed in (11, 12, 13):
            src, _ = random_batch(2, 8, 7, CFG["vs"], CFG["vt"], seed=seed)
            outs = []
            for bs in (bs_graph, bs_loop):
                feed_transformer(model, src, None, train=False)
                bs.reset_batch()
                bs.batch_size = 2
                outs.append(bs.outputs)
            a, b = outs[0].last_search_step_output, outs[1].last_search_step_output
            assert a.token_ids.shape == b.token_ids.shape
            assert bool((a.token_ids[1:] == b.token_ids[1:]).all())
            assert max_abs(a.scores, b.scores) < 1e-4
        assert len(bs_graph._graphs) == 1
    finally:
        ops.set_gemm_backend("auto")
