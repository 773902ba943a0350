"""Golden vectors from the reference's own *pure* functions of the hot path, executed over the numpy
TensorFlow stand-in of tf_numpy_shim.py (see its docstring for what that pins):

  encoders/transformer.py     position_signal
  attention/scaled_dot_product.py  split_for_heads, mask_energies, mask_future, attention (1 and 3 heads)
  tf_utils.py                 layer_norm, gather_flat, partial_transpose, append_tensor
  nn/projection.py            maxout
  functions.py                noam_decay
  decoders/beam_search_decoder.py  BeamSearchDecoder._length_penalty

    python tests/golden/make_tf_shim_golden.py   ->  tests/golden/tf_shim_golden.npz
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim as shim  # noqa: E402
from make_host_golden import install_stubs  # noqa: E402


def main():
    install_stubs()          # termcolor, typeguard, collections aliases, sys.path
    shim.install()           # replaces the empty tensorflow stub
    rng = np.random.RandomState(11)
    out = {}

    def f32(*shape, scale=1.0):
        return np.asarray(rng.randn(*shape) * scale, dtype=np.float32)

    # ---- position signal ---------------------------------------------------------------------
    from neuralmonkey.encoders.transformer import position_signal
    for dim, length in ((6, 7), (512, 50), (9, 4)):
        out["pos_{}_{}".format(dim, length)] = np.asarray(position_signal(dim, length))

    # ---- scaled dot-product attention ------------------------------------------------------------
    from neuralmonkey.attention import scaled_dot_product as sdp
    q, k, v = f32(2, 5, 12), f32(2, 7, 12), f32(2, 7, 12)
    mask = np.array([[1, 1, 1, 1, 1, 0, 0], [1, 1, 1, 0, 0, 0, 0]], np.float32)
    out.update({"att_q": q, "att_k": k, "att_v": v, "att_mask": mask})
    out["split_heads"] = np.asarray(sdp.split_for_heads(shim.t(q), 3, 4))
    energies = f32(2, 3, 5, 7)
    out["energies"] = energies
    out["mask_energies"] = np.asarray(sdp.mask_energies(shim.t(energies), shim.t(mask)))
    sq = f32(2, 3, 5, 5)
    out["energies_sq"] = sq
    out["mask_future"] = np.asarray(sdp.mask_future(shim.t(sq)))
    ctx, weights = sdp.attention(shim.t(q), shim.t(k), shim.t(v), shim.t(mask), 1, lambda x: x)
    out["att1_ctx"], out["att1_w"] = np.asarray(ctx), np.asarray(weights)
    for name in ("query_proj", "keys_proj", "vals_proj", "output_proj"):
        shim.DENSE[name] = (f32(12, 12, scale=0.3), None)
        out["dense_" + name] = shim.DENSE[name][0]
    ctx, weights = sdp.attention(shim.t(q), shim.t(k), shim.t(v), shim.t(mask), 3, lambda x: x)
    out["att3_ctx"], out["att3_w"] = np.asarray(ctx), np.asarray(weights)
    qs = f32(2, 6, 12)
    smask = np.array([[1, 1, 1, 1, 0, 0], [1, 1, 1, 1, 1, 1]], np.float32)
    ctx, weights = sdp.attention(shim.t(qs), shim.t(qs), shim.t(qs), shim.t(smask), 3, lambda x: x, masked=True)
    out.update({"self_q": qs, "self_mask": smask, "self_ctx": np.asarray(ctx), "self_w": np.asarray(weights)})
    # attention-weight dropout: the callback receives the softmax weights; here it applies a FIXED mask
    # (0 or 1/keep_prob), so that where the dropout sits is pinned without a random stream
    drop_rng = np.random.RandomState(3)
    drop = (drop_rng.rand(2, 3, 6, 6) < 0.7).astype(np.float32) / 0.7
    ctx, weights = sdp.attention(shim.t(qs), shim.t(qs), shim.t(qs), shim.t(smask), 3,
                                 lambda w: shim.t(np.asarray(w) * drop), masked=True)
    out.update({"drop_mask": drop, "drop_ctx": np.asarray(ctx), "drop_w": np.asarray(weights)})

    # ---- tf_utils ------------------------------------------------------------------------------------
    from neuralmonkey import tf_utils
    x = f32(4, 5, 10, scale=2.0)
    shim.VARIABLES["LayerNorm/gamma"] = (1.0 + f32(10, scale=0.2))
    shim.VARIABLES["LayerNorm/beta"] = f32(10, scale=0.2)
    out.update({"ln_x": x, "ln_gamma": shim.VARIABLES["LayerNorm/gamma"], "ln_beta": shim.VARIABLES["LayerNorm/beta"]})
    out["ln_y"] = np.asarray(tf_utils.layer_norm(shim.t(x)))
    batch, beam = 3, 4
    state = f32(batch * beam, 5)
    beam_ids = rng.randint(0, beam, size=(batch, beam))
    idx = np.stack([np.tile(np.arange(batch)[:, None], (1, beam)), beam_ids], axis=2)
    out.update({"gf_state": state, "gf_beam_ids": beam_ids.astype(np.int32)})
    out["gf_out"] = np.asarray(tf_utils.gather_flat(shim.t(state), shim.t(idx), batch, beam))
    hist = f32(6, batch * beam, 2)
    out["pt_in"] = hist
    out["pt_out"] = np.asarray(tf_utils.partial_transpose(shim.t(hist), [1, 0]))
    out["append_out"] = np.asarray(tf_utils.append_tensor(shim.t(hist), shim.t(hist[0] * 2), 0))

    # ---- maxout ----------------------------------------------------------------------------------------
    from neuralmonkey.nn.projection import maxout
    inp = f32(5, 8)
    shim.DENSE["MaxoutProjection"] = (f32(8, 6, scale=0.5), f32(6, scale=0.5))
    out.update({"maxout_in": inp, "maxout_kernel": shim.DENSE["MaxoutProjection"][0],
                "maxout_bias": shim.DENSE["MaxoutProjection"][1]})
    out["maxout_out"] = np.asarray(maxout(shim.t(inp), 3))

    # ---- noam decay, length penalty ---------------------------------------------------------------------
    from neuralmonkey.functions import noam_decay
    steps = [0, 1, 50, 111, 112, 400, 4000]
    vals = []
    for step in steps:
        shim.GLOBAL_STEP[0] = step
        with np.errstate(divide="ignore"):
            vals.append(float(np.asarray(noam_decay(0.2, 6, 111))))
    out["noam_steps"], out["noam_values"] = np.array(steps), np.array(vals)
    from neuralmonkey.decoders.beam_search_decoder import BeamSearchDecoder
    lengths = np.arange(0, 40, dtype=np.int32).reshape(4, 10)
    for alpha in (0.0, 0.6, 1.0):
        dummy = types.SimpleNamespace(length_normalization=alpha)
        out["lp_{}".format(alpha)] = np.asarray(BeamSearchDecoder._length_penalty(dummy, shim.t(lengths)))
    out["lp_lengths"] = lengths
    # ---- Bahdanau attention: Attention.attention (attention/feed_forward.py:125-166) -----------------
    from neuralmonkey.attention.feed_forward import Attention
    from neuralmonkey.attention.namedtuples import AttentionLoopState
    states = f32(3, 6, 10)
    amask = np.array([[1, 1, 1, 1, 1, 1], [1, 1, 1, 0, 0, 0], [1, 0, 0, 0, 0, 0]], np.float32)
    query = f32(3, 8)
    for name, shape in (("Attention/attn_query_projection", (8, 7)), ("attn_key_projection", (10, 7)),
                        ("attn_similarity_v", (7,)), ("attn_projection_bias", (7,)), ("attn_bias", ())):
        shim.VARIABLES["att/" + name] = f32(*shape, scale=0.5)
        out["bah_" + name.split("/")[-1]] = shim.VARIABLES["att/" + name]
    out.update({"bah_states": states, "bah_mask": amask, "bah_query": query})
    for label, use_mask in (("masked", True), ("nomask", False)):
        att = object.__new__(Attention)
        att._variable_scope = shim.VarScope("att")
        att._reuse = None
        att._name = "att"
        att._state_size = 7
        att._attention_states_cached_placeholder = shim.t(states)
        att._attention_mask_cached_placeholder = shim.t(amask) if use_mask else None
        empty = AttentionLoopState(contexts=shim.t(np.zeros((0, 3, 10), np.float32)),
                                   weights=shim.t(np.zeros((0, 3, 6), np.float32)))
        ctx, loop_state = att.attention(shim.t(query), None, None, empty)
        out["bah_ctx_" + label] = np.asarray(ctx)
        out["bah_w_" + label] = np.asarray(loop_state.weights)[0]

    # ---- decoder projections (decoders/encoder_projection.py, decoders/output_projection.py) -------------
    from neuralmonkey.decoders.encoder_projection import linear_encoder_projection
    from neuralmonkey.decoders.output_projection import maxout_output, nonlinear_output
    enc_a, enc_b = f32(4, 6), f32(4, 3)
    shim.DENSE["encoders_projection"] = (f32(9, 5, scale=0.4), f32(5, scale=0.4))
    init = linear_encoder_projection(1.0)(None, 5, [types.SimpleNamespace(output=shim.t(enc_a)),
                                                    types.SimpleNamespace(output=shim.t(enc_b))])
    out.update({"proj_enc_a": enc_a, "proj_enc_b": enc_b, "proj_kernel": shim.DENSE["encoders_projection"][0],
                "proj_bias": shim.DENSE["encoders_projection"][1], "proj_init": np.asarray(init)})
    cell, emb, ctx1 = f32(4, 5), f32(4, 3), f32(4, 6)
    shim.DENSE[None] = (f32(14, 4, scale=0.4), f32(4, scale=0.4))          # tf.layers.dense without a name
    fn, size = nonlinear_output(4)
    out.update({"op_cell": cell, "op_emb": emb, "op_ctx": ctx1, "op_tanh_kernel": shim.DENSE[None][0],
                "op_tanh_bias": shim.DENSE[None][1],
                "op_tanh_out": np.asarray(fn(shim.t(cell), shim.t(emb), [shim.t(ctx1)], None))})
    shim.DENSE["MaxoutProjection"] = (f32(14, 8, scale=0.4), f32(8, scale=0.4))
    fn, size = maxout_output(4)
    out.update({"op_max_kernel": shim.DENSE["MaxoutProjection"][0], "op_max_bias": shim.DENSE["MaxoutProjection"][1],
                "op_max_out": np.asarray(fn(shim.t(cell), shim.t(emb), [shim.t(ctx1)], None))})
    # ---- whole Transformer stacks: TransformerEncoder.temporal_states / output and
    #      TransformerDecoder.layer(depth, ...) with every variable looked up by its full TF name ----------
    from neuralmonkey.encoders.transformer import TransformerEncoder
    from neuralmonkey.decoders.transformer import TransformerDecoder
    dim, ff, depth, heads = 12, 20, 2, 3

    def ln_vars(scope):
        shim.VARIABLES[scope + "/LayerNorm/gamma"] = 1.0 + f32(dim, scale=0.2)
        shim.VARIABLES[scope + "/LayerNorm/beta"] = f32(dim, scale=0.2)

    def att_vars(scope):
        for name in ("query_proj", "keys_proj", "vals_proj", "output_proj"):
            shim.VARIABLES["{}/{}/kernel".format(scope, name)] = f32(dim, dim, scale=0.3)

    def ff_vars(scope):
        ln_vars(scope)
        shim.VARIABLES[scope + "/hidden_state/kernel"], shim.VARIABLES[scope + "/hidden_state/bias"] = f32(dim, ff, scale=0.3), f32(ff, scale=0.2)
        shim.VARIABLES[scope + "/output/kernel"], shim.VARIABLES[scope + "/output/bias"] = f32(ff, dim, scale=0.3), f32(dim, scale=0.2)

    first_new = len(shim.VARIABLES)
    for i in range(depth):
        ln_vars("tenc/layer_{}/self_attention".format(i)); att_vars("tenc/layer_{}/self_attention".format(i))
        ff_vars("tenc/layer_{}/feedforward".format(i))
        ln_vars("tdec/layer_{}/self_attention".format(i)); att_vars("tdec/layer_{}/self_attention".format(i))
        ln_vars("tdec/layer_{}/encdec_attention/enc_0".format(i)); att_vars("tdec/layer_{}/encdec_attention/enc_0".format(i))
        ff_vars("tdec/layer_{}/feedforward".format(i))
    ln_vars("tenc")
    ln_vars("tdec")
    for name in list(shim.VARIABLES)[first_new:]:
        out["tv::" + name] = shim.VARIABLES[name]
    enc_in = f32(3, 6, dim)
    enc_mask = np.array([[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 0, 0], [1, 0, 0, 0, 0, 0]], np.float32)
    enc = object.__new__(TransformerEncoder)
    enc.__dict__.update(dict(
        input_sequence=types.SimpleNamespace(temporal_states=shim.t(enc_in), temporal_mask=shim.t(enc_mask), dimension=dim),
        ff_hidden_size=ff, depth=depth, n_heads=heads, dropout_keep_prob=1.0, attention_dropout_keep_prob=1.0,
        target_space_id=None, use_att_transform_bias=False, use_positional_encoding=True,
        input_for_cross_attention=None, n_cross_att_heads=None, train_mode=None,
        _variable_scope=shim.VarScope("tenc"), _reuse=None, _name="tenc"))
    shim.USED[:] = []
    enc_states = np.asarray(enc.temporal_states)
    out.update({"tenc_in": enc_in, "tenc_mask": enc_mask, "tenc_states": enc_states,
                "tenc_output": np.asarray(enc.output)})
    dec = object.__new__(TransformerDecoder)
    dec.__dict__.update(dict(
        encoders=[enc], ff_hidden_size=ff, n_heads_self=heads, n_heads_enc=[2 if dim % 2 == 0 else heads], depth=depth,
        attention_dropout_keep_prob=[1.0], self_att_dropout_keep_prob=1.0, dropout_keep_prob=1.0,
        use_att_transform_bias=False, attention_combination_strategy="serial", n_heads_hier=None,
        encoder_states=lambda: [shim.t(enc_states)], encoder_masks=lambda: [shim.t(enc_mask)],
        _embedding_size=dim, embeddings_source=None, train_mode=None,
        _variable_scope=shim.VarScope("tdec"), _reuse=None, _name="tdec"))
    dec_in = f32(3, 5, dim)
    dec_mask = np.array([[1, 1, 1, 1, 1], [1, 1, 1, 0, 0], [1, 1, 0, 0, 0]], np.float32)
    with dec.use_scope():
        last = dec.layer(depth, shim.t(dec_in), shim.t(dec_mask))
    out.update({"tdec_in": dec_in, "tdec_mask": dec_mask, "tdec_states": np.asarray(last.temporal_states)})
    out["transformer_dense_names"] = np.array(sorted(set(shim.USED)))

    # ---- the Transformer decoder's own loops: train_loop_result (decoders/transformer.py:393-453), the
    #      run-time decoding_loop with next_state (:487-516) re-running the prefix, and - further below -
    #      a whole beam search around it (three sentences, encoder states tiled to the beam by
    #      BeamSearchDecoder.outputs) ------------------------------------------------------------------------------
    tvsz, tmax = 11, 6
    ttable = f32(tvsz, dim, scale=0.6)
    tgold = np.array([[5, 6, 2, 0, 0], [7, 8, 9, 4, 2], [4, 2, 0, 0, 0]], np.int64).T
    dec.__dict__.update(dict(
        vocabulary=list(range(tvsz)), supress_unk=False, max_output_len=tmax, batch_size=3, tie_embeddings=True,
        _embedding_matrix_cached_placeholder=shim.t(ttable),
        _go_symbols_cached_placeholder=shim.t(np.full((3,), 1, np.int64)),
        _train_inputs_cached_placeholder=shim.t(tgold)))
    out.update({"tloop_table": ttable.copy(), "tloop_gold": tgold})
    train_ls = dec.train_loop_result
    out["tloop_train_logits"] = np.asarray(train_ls.histories.logits)
    out["tloop_train_states"] = np.asarray(train_ls.histories.output_states)
    out["tloop_train_input_symbols"] = np.asarray(dec.train_input_symbols)
    # supress_unk: the training pass computes its logits itself (:409-419) and never adds the -1e9 <unk>
    # column that get_body's state_to_logits adds at run time (autoregressive.py:454-457)
    unk_dec = object.__new__(TransformerDecoder)
    unk_dec.__dict__.update({k: v for k, v in dec.__dict__.items() if not k.endswith("_cached_placeholder")
                             or k in ("_embedding_matrix_cached_placeholder", "_go_symbols_cached_placeholder",
                                      "_train_inputs_cached_placeholder")})
    unk_dec.supress_unk = True
    out["tloop_unk_train_logits"] = np.asarray(unk_dec.train_loop_result.histories.logits)
    unk_run = unk_dec.decoding_loop(train_mode=False)
    out["tloop_unk_run_logits"] = np.asarray(unk_run.histories.logits)
    run_ls = dec.decoding_loop(train_mode=False)
    out["tloop_run_logits"] = np.asarray(run_ls.histories.logits)
    out["tloop_run_symbols"] = np.asarray(run_ls.histories.output_symbols)
    out["tloop_run_mask"] = np.asarray(run_ls.histories.output_mask)
    out["tloop_run_input_mask"] = np.asarray(run_ls.feedables.other.input_mask)
    transformer_parent = dec
    # ---- one beam search step: BeamSearchDecoder.get_body()() (decoders/beam_search_decoder.py:385-558)
    #      around a stand-in parent decoder whose body returns given logits -----------------------------------
    from neuralmonkey.decoders import beam_search_decoder as bsd
    from neuralmonkey.decoders.autoregressive import (DecoderConstants, DecoderFeedables, DecoderHistories,
                                                      LoopState)
    bsz, beam, vocab_size, edim = 3, 4, 9, 5
    rows = bsz * beam
    emb_table = f32(vocab_size, edim)
    next_logits = f32(rows, vocab_size, scale=2.0)

    def decoder_body(*args):
        ls = LoopState(*args)
        hist = ls.histories._replace(logits=shim.t(np.concatenate(
            [np.asarray(ls.histories.logits), next_logits[None]], 0)))
        return LoopState(histories=hist, constants=ls.constants,
                         feedables=ls.feedables._replace(step=ls.feedables.step + 1))

    parent = types.SimpleNamespace(get_body=lambda train_mode: decoder_body, vocabulary=list(range(vocab_size)),
                                   embed_input_symbols=lambda ids: shim.t(emb_table[np.asarray(ids)]))
    for case, alpha in (("a", 0.6), ("b", 1.0)):
        dec = object.__new__(bsd.BeamSearchDecoder)
        dec.__dict__.update(dict(parent_decoder=parent, beam_size=beam, batch_size=bsz, length_normalization=alpha,
                                 max_steps_int=10, _variable_scope=shim.VarScope("bs"), _reuse=None, _name="bs"))
        prev_logprobs = np.log(np.asarray(shim._softmax(f32(bsz, beam, vocab_size, scale=2.0))))
        if case == "b":                              # exact ties inside a sentence: top_k keeps the lower index
            prev_logprobs[0, 1] = prev_logprobs[0, 0]
        logprob_sum = -np.abs(f32(bsz, beam))
        if case == "b":
            logprob_sum[0, 1] = logprob_sum[0, 0]
        lengths = rng.randint(0, 6, size=(bsz, beam)).astype(np.int32)
        if case == "b":
            lengths[0, 1] = lengths[0, 0]
        finished = np.array([[0, 1, 0, 0], [0, 0, 0, 1], [1, 1, 0, 0]], bool)
        token_ids = rng.randint(4, vocab_size, size=(3, bsz, beam)).astype(np.int64)
        state_feed = f32(rows, 6)
        feedables = DecoderFeedables(step=shim.t(np.int32(3)), finished=shim.t(finished.reshape(-1)),
                                     embedded_input=shim.t(f32(rows, edim)), other=[shim.t(state_feed)])
        histories = DecoderHistories(logits=shim.t(f32(3, rows, vocab_size)), output_states=shim.t(f32(3, rows, 2)),
                                     output_symbols=shim.t(np.zeros((3, rows), np.int64)),
                                     output_mask=shim.t(np.ones((3, rows), bool)), other=[])
        loop_state = bsd.BeamSearchLoopState(
            search_state=bsd.SearchState(logprob_sum=shim.t(logprob_sum), prev_logprobs=shim.t(prev_logprobs),
                                         lengths=shim.t(lengths), finished=shim.t(finished)),
            search_results=bsd.SearchResults(scores=shim.t(np.zeros((bsz, beam), np.float32)),
                                             token_ids=shim.t(token_ids)),
            decoder_loop_state=LoopState(histories=histories, constants=DecoderConstants(train_inputs=shim.t(np.zeros(1))),
                                         feedables=feedables))
        nxt = dec.get_body()(*loop_state)
        pre = "beam_{}_".format(case)
        out.update({pre + "alpha": np.float32(alpha), pre + "prev_logprobs": prev_logprobs, pre + "logprob_sum": logprob_sum,
                    pre + "lengths": lengths, pre + "finished": finished, pre + "token_ids": token_ids,
                    pre + "state_feed": state_feed, pre + "next_logits": next_logits, pre + "emb_table": emb_table,
                    pre + "out_scores": np.asarray(nxt.search_results.scores),
                    pre + "out_token_ids": np.asarray(nxt.search_results.token_ids),
                    pre + "out_logprob_sum": np.asarray(nxt.search_state.logprob_sum),
                    pre + "out_lengths": np.asarray(nxt.search_state.lengths),
                    pre + "out_finished": np.asarray(nxt.search_state.finished),
                    pre + "out_prev_logprobs": np.asarray(nxt.search_state.prev_logprobs),
                    pre + "out_state_feed": np.asarray(nxt.decoder_loop_state.feedables.other[0]),
                    pre + "out_embedded": np.asarray(nxt.decoder_loop_state.feedables.embedded_input),
                    pre + "out_dec_finished": np.asarray(nxt.decoder_loop_state.feedables.finished)})
    # ---- the decoding loop every decoder shares: AutoregressiveDecoder.get_initial_loop_state /
    #      loop_continue_criterion / get_body (decoders/autoregressive.py:381-519), driven by a python
    #      `while` in place of tf.while_loop, around a stand-in next_state ------------------------------------
    from neuralmonkey.decoders.autoregressive import AutoregressiveDecoder
    vsz, odim, edim2, max_len, nb = 11, 6, 6, 7, 5   # output size == embedding size, as the reference requires
    dec_w, dec_b, table = f32(odim, vsz), f32(vsz, scale=0.3), f32(vsz, edim2)
    step_states = f32(max_len, nb, odim)
    gold = np.array([[5, 6, 2, 0, 0, 0], [7, 8, 9, 4, 2, 0], [4, 2, 0, 0, 0, 0], [3, 3, 3, 3, 3, 2], [6, 2, 0, 0, 0, 0]],
                    np.int64).T                     # time-major [T, B], incl. </s>, padded
    # make the runtime argmax hit </s> at different steps for different sentences
    dec_b_run = dec_b.copy()
    out.update({"loop_w": dec_w, "loop_b": dec_b, "loop_table": table, "loop_states": step_states, "loop_gold": gold})
    for mode, supress, eos_bonus in (("train", True, 0.0), ("run", False, 0.0), ("run_unk", True, 0.0),
                                     ("run_eos", True, 2.5)):
        bias = dec_b.copy()
        bias[2] += eos_bonus
        out["loop_{}_bias".format(mode)] = bias
        ad = object.__new__(AutoregressiveDecoder)
        ad.__dict__.update(dict(
            vocabulary=list(range(vsz)), supress_unk=supress, max_output_len=max_len, batch_size=nb,
            dropout_keep_prob=1.0, train_mode=None, _embedding_size=edim2, embeddings_source=None,
            _variable_scope=shim.VarScope("dec"), _reuse=None, _name="dec",
            _decoding_w_cached_placeholder=shim.t(dec_w), _decoding_b_cached_placeholder=shim.t(bias),
            _embedding_matrix_cached_placeholder=shim.t(table),
            _go_symbols_cached_placeholder=shim.t(np.full((nb,), 1, np.int64)),
            _train_inputs_cached_placeholder=shim.t(gold)))
        ad.next_state = lambda ls: (shim.t(step_states[int(ls.feedables.step)]), None, None)
        body = ad.get_body(train_mode=(mode == "train"))
        state = ad.get_initial_loop_state()
        while bool(np.asarray(ad.loop_continue_criterion(*state))):
            state = body(*state)
        out["loop_{}_logits".format(mode)] = np.asarray(state.histories.logits)
        out["loop_{}_symbols".format(mode)] = np.asarray(state.histories.output_symbols)
        out["loop_{}_mask".format(mode)] = np.asarray(state.histories.output_mask)
        out["loop_{}_steps".format(mode)] = np.int64(np.asarray(state.feedables.step))
        out["loop_{}_last_input".format(mode)] = np.asarray(state.feedables.embedded_input)

    # ---- the whole attention decoder: decoders/decoder.py Decoder (initial_state, get_initial_feedables /
    #      histories, next_state) under AutoregressiveDecoder's loop, with the reference's own Attention,
    #      encoder projection and output projections; only the GRU cell arithmetic is the shim's
    #      restatement of TensorFlow's published GRUCell.  Variables are looked up by their full TF names.
    from neuralmonkey.decoders.decoder import Decoder
    from neuralmonkey.decoders.autoregressive import LoopState

    built = {}

    def rnn_decoder_case(tag, hsz, esz, maxout, use_mask, nb=4, modes=("train", "run", "loss"), eos_bonus=0.0,
                         short_gold=False):
        tx, csz, asz, vsz, max_len = 6, 10, 8, 13, 6
        dname, aname = "rd_" + tag, "ra_" + tag
        first = len(shim.VARIABLES)
        var = shim.VARIABLES
        var[aname + "/Attention/attn_query_projection"] = f32(hsz, asz, scale=0.5)
        var[aname + "/attn_key_projection"] = f32(csz, asz, scale=0.5)
        var[aname + "/attn_similarity_v"] = f32(asz, scale=0.7)
        var[aname + "/attn_projection_bias"] = f32(asz, scale=0.3)
        var[aname + "/attn_bias"] = f32(scale=0.3)
        var[dname + "/initial_state/encoders_projection/kernel"] = f32(csz, hsz, scale=0.4)
        var[dname + "/initial_state/encoders_projection/bias"] = f32(hsz, scale=0.3)
        cell = dname + "/attention_decoder/OrthoGRUCell/"
        var[cell + "gates/kernel"], var[cell + "gates/bias"] = f32(esz + hsz, 2 * hsz, scale=0.5), 1.0 + f32(2 * hsz, scale=0.2)
        var[cell + "candidate/kernel"], var[cell + "candidate/bias"] = f32(esz + hsz, hsz, scale=0.5), f32(hsz, scale=0.2)
        if maxout:
            proj = dname + "/attention_decoder/MaxoutProjection/MaxoutProjection/"
            var[proj + "kernel"], var[proj + "bias"] = f32(hsz + esz + csz, 2 * esz, scale=0.4), f32(2 * esz, scale=0.3)
        else:
            proj = dname + "/attention_decoder/dense/"
            var[proj + "kernel"], var[proj + "bias"] = f32(hsz + esz + csz, hsz, scale=0.4), f32(hsz, scale=0.3)
        for name in list(var)[first:]:
            out["rv::" + name] = var[name]
        states, enc_out = f32(nb, tx, csz), f32(nb, csz)
        amask = np.array([[1, 1, 1, 1, 0, 0], [1, 1, 1, 1, 1, 1], [1, 1, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0]], np.float32)[:nb]
        dec_w, dec_b, table = f32(esz, vsz, scale=0.8), f32(vsz, scale=0.3), f32(vsz, esz)
        dec_b[2] += eos_bonus
        gold = np.array([[5, 6, 7, 2, 0, 0], [7, 8, 9, 4, 3, 2], [4, 2, 0, 0, 0, 0], [3, 11, 12, 10, 2, 0]], np.int64)[:nb].T
        if short_gold:                       # training ends after 4 steps, the run-time loop goes on: min_time
            gold = np.array([[5, 6, 7, 2], [7, 8, 2, 0], [4, 2, 0, 0], [3, 11, 12, 2]], np.int64)[:nb].T
        out.update({"rd_{}_states".format(tag): states, "rd_{}_enc_out".format(tag): enc_out,
                    "rd_{}_mask".format(tag): amask, "rd_{}_w".format(tag): dec_w, "rd_{}_b".format(tag): dec_b,
                    "rd_{}_table".format(tag): table, "rd_{}_gold".format(tag): gold})
        for mode in modes:
            att = object.__new__(Attention)
            att.__dict__.update(dict(
                _variable_scope=shim.VarScope(aname), _reuse=None, _name=aname, _state_size=asz, batch_size=nb,
                _histories={}, _attention_states_cached_placeholder=shim.t(states),
                _attention_mask_cached_placeholder=shim.t(amask) if use_mask else None))
            rd = object.__new__(Decoder)
            rd.__dict__.update(dict(
                vocabulary=list(range(vsz)), supress_unk=False, max_output_len=max_len, batch_size=nb,
                label_smoothing=None,
                dropout_keep_prob=1.0, train_mode=None, _embedding_size=esz, embeddings_source=None,
                _variable_scope=shim.VarScope(dname), _reuse=None, _name=dname,
                _decoding_w_cached_placeholder=shim.t(dec_w), _decoding_b_cached_placeholder=shim.t(dec_b),
                _embedding_matrix_cached_placeholder=shim.t(table),
                _go_symbols_cached_placeholder=shim.t(np.full((nb,), 1, np.int64)),
                _train_inputs_cached_placeholder=shim.t(gold),
                encoders=[types.SimpleNamespace(output=shim.t(enc_out))],
                _output_projection_spec=maxout_output(esz) if maxout else None, _conditional_gru=False,
                _attention_on_input=False, _rnn_cell_str="GRU", _rnn_size=hsz, _encoder_projection=None,
                attentions=[att], step_scope=shim.VarScope(dname + "/attention_decoder"),
                encoder_states=lambda: [], encoder_masks=lambda: [],
                input_projection=lambda *args: LoopState(*args).feedables.embedded_input))
            built[tag] = rd
            if mode == "build":
                return
            if mode == "loss":                   # the lazily built loss tensors (autoregressive.py:292-371)
                key = "rd_{}_".format(tag)
                # label smoothing (:294-300): tf.losses.softmax_cross_entropy reduces to ONE scalar, which
                # sequence_loss then multiplies by the weights
                smooth = object.__new__(Decoder)
                smooth.__dict__.update(rd.__dict__)
                smooth.__dict__.pop("_train_xents_cached_placeholder", None)
                smooth.__dict__.pop("_train_loss_cached_placeholder", None)
                smooth.label_smoothing = 0.1
                smooth._train_loop_result_cached_placeholder = rd.train_loop_result
                smooth._train_logits_cached_placeholder = rd.train_logits
                out[key + "smooth_train_xents"] = np.asarray(smooth.train_xents)
                out[key + "smooth_train_loss"] = np.asarray(smooth.train_loss)
                for attr in ("train_xents", "train_loss", "train_mask", "runtime_xents", "runtime_loss", "decoded",
                             "runtime_logprobs", "runtime_mask"):
                    out[key + attr] = np.asarray(getattr(rd, attr))
                continue
            shim.USED[:] = []
            shim.GRUCell.CALLS[:] = []
            body = rd.get_body(train_mode=(mode == "train"))
            state = rd.get_initial_loop_state()
            while bool(np.asarray(rd.loop_continue_criterion(*state))):
                state = body(*state)
            rd.finalize_loop(state, mode == "train")
            key = "rd_{}_{}_".format(tag, mode)
            out[key + "logits"] = np.asarray(state.histories.logits)
            out[key + "output_states"] = np.asarray(state.histories.output_states)
            out[key + "symbols"] = np.asarray(state.histories.output_symbols)
            out[key + "out_mask"] = np.asarray(state.histories.output_mask)
            out[key + "rnn_outputs"] = np.asarray(state.histories.other.rnn_outputs)
            out[key + "att_weights"] = np.asarray(att.histories["{}_{}".format(dname, mode)])
            out[key + "att_contexts"] = np.asarray(state.histories.other.attention_histories[0].contexts)
            out[key + "initial_state"] = np.asarray(rd.initial_state)
        out["rd_{}_dense_names".format(tag)] = np.array(sorted(set(shim.USED)))
        out["rd_{}_cell_scopes".format(tag)] = np.array(sorted({c[0] for c in shim.GRUCell.CALLS}))

    rnn_decoder_case("maxout", hsz=7, esz=5, maxout=True, use_mask=True)
    rnn_decoder_case("tanh", hsz=6, esz=6, maxout=False, use_mask=False, short_gold=True)

    # ---- decoder variants: NematusGRUCell (the reference's own cell code, nn/ortho_gru_cell.py:
    #      57-105), the conditional GRU (decoder.py:303-325), attention_on_input (:264-277), nematus_output /
    #      mlp_output (output_projection.py:76-112,163-188), nematus_projection / concat / empty initial states
    #      (encoder_projection.py:30-145) ------------------------------------------------------------------------
    from neuralmonkey.decoders.output_projection import nematus_output, mlp_output
    from neuralmonkey.decoders.encoder_projection import (nematus_projection, concat_encoder_projection,
                                                          empty_initial_state)

    def rnn_variant_case(tag, cell, conditional, att_on_input, out_proj, enc_proj, hsz=7, esz=5):
        nb, tx, csz, asz, vsz, max_len = 3, 5, 10, 8, 12, 5
        if enc_proj == "concat":
            hsz = csz
        dname, aname = "vd_" + tag, "va_" + tag
        first = len(shim.VARIABLES)
        var = shim.VARIABLES
        var[aname + "/Attention/attn_query_projection"] = f32(hsz, asz, scale=0.5)
        var[aname + "/attn_key_projection"] = f32(csz, asz, scale=0.5)
        var[aname + "/attn_similarity_v"] = f32(asz, scale=0.7)
        var[aname + "/attn_projection_bias"] = f32(asz, scale=0.3)
        var[aname + "/attn_bias"] = f32(scale=0.3)
        if enc_proj in ("linear", "nematus"):
            var[dname + "/initial_state/encoders_projection/kernel"] = f32(csz, hsz, scale=0.4)
            var[dname + "/initial_state/encoders_projection/bias"] = f32(hsz, scale=0.3)
        step = dname + "/attention_decoder/"

        def nematus_cell(scope, in_dim, state_bias, input_bias):
            for part, width in (("gates", 2 * hsz), ("candidate", hsz)):
                var[scope + part + "/input_proj/kernel"] = f32(in_dim, width, scale=0.5)
                var[scope + part + "/state_proj/kernel"] = f32(hsz, width, scale=0.5)
                if input_bias:
                    var[scope + part + "/input_proj/bias"] = f32(width, scale=0.3)
                if state_bias:
                    var[scope + part + "/state_proj/bias"] = f32(width, scale=0.3)

        def tf_cell(scope, in_dim):
            var[scope + "gates/kernel"], var[scope + "gates/bias"] = f32(in_dim + hsz, 2 * hsz, scale=0.5), 1.0 + f32(2 * hsz, scale=0.2)
            var[scope + "candidate/kernel"], var[scope + "candidate/bias"] = f32(in_dim + hsz, hsz, scale=0.5), f32(hsz, scale=0.2)

        if cell == "LSTM":
            var[step + "lstm_cell/kernel"], var[step + "lstm_cell/bias"] = f32(esz + hsz, 4 * hsz, scale=0.5), f32(4 * hsz, scale=0.3)
        elif cell == "NematusGRU":
            nematus_cell(step + "nematus_gru_cell/", esz, False, True)
            if conditional:
                nematus_cell(step + "cond_gru_2_cell/", csz, True, False)
        else:
            tf_cell(step + "OrthoGRUCell/", esz)
            if conditional:
                tf_cell(step + "cond_gru_2_cell/", csz)
        if att_on_input:
            var[step + "dense/kernel"], var[step + "dense/bias"] = f32(esz + csz, esz, scale=0.4), f32(esz, scale=0.3)
        cat = hsz + esz + csz
        if out_proj == "nematus":
            for name, width in (("rnn_state", hsz), ("prev_out", esz), ("context", csz)):
                var[step + name + "/kernel"], var[step + name + "/bias"] = f32(width, esz, scale=0.4), f32(esz, scale=0.3)
            spec_out = nematus_output(esz)
        elif out_proj == "mlp":
            var[step + "deep_output_mlp/mlp_layer_0/kernel"], var[step + "deep_output_mlp/mlp_layer_0/bias"] = f32(cat, 9, scale=0.4), f32(9, scale=0.3)
            var[step + "deep_output_mlp/mlp_layer_1/kernel"], var[step + "deep_output_mlp/mlp_layer_1/bias"] = f32(9, esz, scale=0.4), f32(esz, scale=0.3)
            spec_out = mlp_output([9, esz])
        else:
            var[step + "MaxoutProjection/MaxoutProjection/kernel"] = f32(cat, 2 * esz, scale=0.4)
            var[step + "MaxoutProjection/MaxoutProjection/bias"] = f32(2 * esz, scale=0.3)
            spec_out = maxout_output(esz)
        for name in list(var)[first:]:
            out["vv::" + name] = var[name]
        states, enc_out = f32(nb, tx, csz), f32(nb, csz)
        amask = np.array([[1, 1, 1, 1, 0], [1, 1, 1, 1, 1], [1, 1, 0, 0, 0]], np.float32)
        dec_w, dec_b, table = f32(esz, vsz, scale=0.8), f32(vsz, scale=0.3), f32(vsz, esz)
        gold = np.array([[5, 6, 7, 2, 0], [7, 8, 9, 4, 2], [4, 2, 0, 0, 0]], np.int64).T
        pre = "vd_{}_".format(tag)
        out.update({pre + "states": states, pre + "enc_out": enc_out, pre + "mask": amask, pre + "w": dec_w,
                    pre + "b": dec_b, pre + "table": table, pre + "gold": gold})
        att = object.__new__(Attention)
        att.__dict__.update(dict(
            _variable_scope=shim.VarScope(aname), _reuse=None, _name=aname, _state_size=asz, batch_size=nb,
            _histories={}, _attention_states_cached_placeholder=shim.t(states),
            _attention_mask_cached_placeholder=shim.t(amask), context_vector_size=csz))
        encoder = types.SimpleNamespace(output=shim.t(enc_out), temporal_states=shim.t(states),
                                        temporal_mask=shim.t(amask))
        rd = object.__new__(Decoder)
        rd.__dict__.update(dict(
            vocabulary=list(range(vsz)), supress_unk=False, max_output_len=max_len, batch_size=nb, label_smoothing=None,
            dropout_keep_prob=1.0, train_mode=None, _embedding_size=esz, embeddings_source=None,
            _variable_scope=shim.VarScope(dname), _reuse=None, _name=dname,
            _decoding_w_cached_placeholder=shim.t(dec_w), _decoding_b_cached_placeholder=shim.t(dec_b),
            _embedding_matrix_cached_placeholder=shim.t(table),
            _go_symbols_cached_placeholder=shim.t(np.full((nb,), 1, np.int64)),
            _train_inputs_cached_placeholder=shim.t(gold), encoders=[] if enc_proj == "empty" else [encoder],
            _output_projection_spec=spec_out, _conditional_gru=conditional, _attention_on_input=att_on_input,
            _rnn_cell_str=cell, _rnn_size=None if enc_proj == "concat" else hsz,
            _encoder_projection={"linear": None, "nematus": nematus_projection(), "concat": None,
                                 "empty": None}[enc_proj],
            attentions=[att], step_scope=shim.VarScope(dname + "/attention_decoder"),
            encoder_states=lambda: [], encoder_masks=lambda: []))
        rd.input_projection = (rd.input_plus_attention if att_on_input
                               else (lambda *args: LoopState(*args).feedables.embedded_input))
        shim.USED[:] = []
        shim.GRUCell.CALLS[:] = []
        out[pre + "initial_state"] = np.asarray(rd.initial_state)
        out[pre + "train_logits"] = np.asarray(rd.train_logits)
        out[pre + "train_rnn_outputs"] = np.asarray(rd.train_loop_result.histories.other.rnn_outputs)
        out[pre + "train_loss"] = np.asarray(rd.train_loss)
        out[pre + "run_logits"] = np.asarray(rd.runtime_logits)
        out[pre + "run_symbols"] = np.asarray(rd.runtime_loop_result.histories.output_symbols)
        out[pre + "dense_names"] = np.array(sorted(set(shim.USED)))
        out[pre + "cell_scopes"] = np.array(sorted({c[0] for c in shim.GRUCell.CALLS}))

    rnn_variant_case("nematus", "NematusGRU", True, False, "nematus", "nematus")
    rnn_variant_case("cond_gru", "GRU", True, False, "mlp", "concat")
    rnn_variant_case("nematus_plain", "NematusGRU", False, False, "maxout", "empty")
    rnn_variant_case("lstm", "LSTM", False, False, "maxout", "linear")
    # attention_on_input=True cannot be built in the reference: input_plus_attention reads
    # `feedables.prev_contexts` (decoder.py:273), which lives in `feedables.other` -> AttributeError
    try:
        rnn_variant_case("input_feeding", "GRU", False, True, "maxout", "linear")
        out["attention_on_input_error"] = np.array("")
    except AttributeError as exc:
        out["attention_on_input_error"] = np.array(str(exc))
        for key in [k for k in out if k.startswith("vd_input_feeding_") or k.startswith("vv::vd_input_feeding")
                    or k.startswith("vv::va_input_feeding")]:
            del out[key]

    # ---- a whole beam search: BeamSearchDecoder.get_initial_loop_state / loop_continue_criterion / get_body
    #      (decoders/beam_search_decoder.py:218-558) around the reference's own attention Decoder, one
    #      sentence (the reference's RNN decoder does not tile the encoder states to the beam) --------------
    def beam_search_case(tag, parent, bsz, beam, alpha, max_steps):
        bs = object.__new__(bsd.BeamSearchDecoder)
        bs.__dict__.update(dict(parent_decoder=parent, beam_size=beam, batch_size=bsz, length_normalization=alpha,
                                max_steps_int=max_steps, max_steps=max_steps, _initial_loop_state=None,
                                _variable_scope=shim.VarScope("bs_" + tag), _reuse=None, _name="bs_" + tag))
        result = bs.outputs               # the reference's own driver: initial state + tf.while_loop
        state = bs.initial_loop_state
        pre = "bsearch_{}_".format(tag)
        out[pre + "init_token_ids"] = np.asarray(state.search_results.token_ids)
        out[pre + "init_prev_logprobs"] = np.asarray(state.search_state.prev_logprobs)
        out[pre + "init_logprob_sum"] = np.asarray(state.search_state.logprob_sum)
        state = types.SimpleNamespace(search_results=result.last_search_step_output,
                                      search_state=result.last_search_state)
        steps = np.asarray(state.search_results.token_ids).shape[0] - 1
        out.update({pre + "steps": np.int64(steps), pre + "alpha": np.float32(alpha), pre + "beam": np.int64(beam),
                    pre + "max_steps": np.int64(max_steps),
                    pre + "scores": np.asarray(state.search_results.scores),
                    pre + "token_ids": np.asarray(state.search_results.token_ids),
                    pre + "logprob_sum": np.asarray(state.search_state.logprob_sum),
                    pre + "lengths": np.asarray(state.search_state.lengths),
                    pre + "finished": np.asarray(state.search_state.finished)})

    beam_search_case("tr", transformer_parent, bsz=3, beam=3, alpha=0.6, max_steps=5)
    ttable[2] = ttable[2] * 2.5              # tied embeddings: a long </s> row makes hypotheses finish early
    transformer_parent._embedding_matrix_cached_placeholder = shim.t(ttable)
    del transformer_parent.__dict__["_decoding_w_cached_placeholder"]
    out["tloop_table_eos"] = ttable.copy()
    beam_search_case("tr_eos", transformer_parent, bsz=3, beam=4, alpha=1.0, max_steps=6)
    run_ls = transformer_parent.decoding_loop(train_mode=False)
    out["tloop_run_eos_logits"] = np.asarray(run_ls.histories.logits)
    out["tloop_run_eos_symbols"] = np.asarray(run_ls.histories.output_symbols)
    out["tloop_run_eos_mask"] = np.asarray(run_ls.histories.output_mask)
    out["tloop_run_eos_input_mask"] = np.asarray(run_ls.feedables.other.input_mask)
    rnn_decoder_case("beam", hsz=7, esz=5, maxout=True, use_mask=True, nb=1, modes=("build",))
    beam_search_case("rnn", built["beam"], bsz=1, beam=3, alpha=0.6, max_steps=6)
    rnn_decoder_case("beam1", hsz=6, esz=6, maxout=False, use_mask=True, nb=1, modes=("build",), eos_bonus=2.0)
    beam_search_case("rnn1", built["beam1"], bsz=1, beam=4, alpha=1.0, max_steps=5)
    rnn_decoder_case("beam2", hsz=6, esz=6, maxout=False, use_mask=False, nb=1, modes=("build",), eos_bonus=3.5)
    beam_search_case("rnn2", built["beam2"], bsz=1, beam=2, alpha=0.0, max_steps=12)

    # ---- the recurrent encoder: model/sequence.py EmbeddedFactorSequence.temporal_states / temporal_mask and
    #      encoders/recurrent.py RecurrentEncoder.rnn + rnn_layer, over the shim's restatement of
    #      tf.nn.(bidirectional_)dynamic_rnn and GRUCell (TensorFlow library code) ----------------------------
    from neuralmonkey.model.sequence import EmbeddedFactorSequence
    from neuralmonkey.encoders.recurrent import RecurrentEncoder, _make_rnn_spec

    def recurrent_case(tag, emb_sizes, layers, residual, layer_norm_, final_norm, scale):
        name = "re_" + tag
        first = len(shim.VARIABLES)
        var = shim.VARIABLES
        vocab_sizes = [9, 5][:len(emb_sizes)]
        for i, (vs_, es_) in enumerate(zip(vocab_sizes, emb_sizes)):
            var["{}_input/embedding_matrix_{}".format(name, i)] = f32(vs_, es_)
        in_dim = sum(emb_sizes)
        for i, layer in enumerate(layers):
            size, direction = layer[0], layer[1]
            scope = "{}/rnn_{}_{}".format(name, i, direction)
            ctype = layers[i][2] if len(layers[i]) > 2 else "GRU"
            cname = {"NematusGRU": "nematus_gru_cell/", "LSTM": "lstm_cell/", "GRU": "OrthoGRUCell/"}[ctype]
            cells = ([scope + "/bidirectional_rnn/fw/" + cname, scope + "/bidirectional_rnn/bw/" + cname]
                     if direction == "bidirectional" else [scope + "/rnn/" + cname])
            for cell in cells:
                if cname == "lstm_cell/":
                    var[cell + "kernel"], var[cell + "bias"] = f32(in_dim + size, 4 * size, scale=0.5), f32(4 * size, scale=0.3)
                    continue
                if cname == "nematus_gru_cell/":        # RNN_CELL_TYPES["NematusGRU"](size): input bias only
                    for part, width in (("gates", 2 * size), ("candidate", size)):
                        var[cell + part + "/input_proj/kernel"] = f32(in_dim, width, scale=0.5)
                        var[cell + part + "/input_proj/bias"] = f32(width, scale=0.3)
                        var[cell + part + "/state_proj/kernel"] = f32(size, width, scale=0.5)
                    continue
                var[cell + "gates/kernel"], var[cell + "gates/bias"] = f32(in_dim + size, 2 * size, scale=0.5), 1.0 + f32(2 * size, scale=0.2)
                var[cell + "candidate/kernel"], var[cell + "candidate/bias"] = f32(in_dim + size, size, scale=0.5), f32(size, scale=0.2)
            if layer_norm_:
                var[scope + "/LayerNorm/gamma"], var[scope + "/LayerNorm/beta"] = 1.0 + f32(in_dim, scale=0.2), f32(in_dim, scale=0.2)
            in_dim = 2 * size if direction == "bidirectional" else size
        if final_norm:
            var[name + "/LayerNorm/gamma"], var[name + "/LayerNorm/beta"] = 1.0 + f32(in_dim, scale=0.2), f32(in_dim, scale=0.2)
        for vname in list(var)[first:]:
            out["ev::" + vname] = var[vname]
        ids = np.array([[4, 5, 6, 7, 8, 3], [3, 4, 5, 0, 0, 0], [8, 0, 0, 0, 0, 0], [5, 5, 4, 4, 3, 0]], np.int64)
        factors = [ids, (ids % 4 + 1) * (ids != 0)][:len(emb_sizes)]
        seq = object.__new__(EmbeddedFactorSequence)
        seq.__dict__.update(dict(
            _variable_scope=shim.VarScope(name + "_input"), _reuse=None, _name=name + "_input",
            vocabularies=[list(range(v)) for v in vocab_sizes], vocabulary_sizes=vocab_sizes,
            data_ids=["f{}".format(i) for i in range(len(emb_sizes))], embedding_sizes=list(emb_sizes),
            scale_embeddings_by_depth=scale, embeddings_source=None, trainable=True,
            _input_factor_indices_cached_placeholder=[shim.t(f) for f in factors]))
        enc = object.__new__(RecurrentEncoder)
        enc.__dict__.update(dict(
            _variable_scope=shim.VarScope(name), _reuse=None, _name=name, input_sequence=seq,
            dropout_keep_prob=1.0, train_mode=None, rnn_specs=[_make_rnn_spec(*l) for l in layers],
            add_residual=residual, add_layer_norm=layer_norm_, include_final_layer_norm=final_norm))
        shim.GRUCell.CALLS[:] = []
        out.update({name + "_ids": np.stack(factors), name + "_embedded": np.asarray(seq.temporal_states),
                    name + "_mask": np.asarray(seq.temporal_mask), name + "_lengths": np.asarray(seq.lengths),
                    name + "_states": np.asarray(enc.temporal_states), name + "_output": np.asarray(enc.output),
                    name + "_enc_mask": np.asarray(enc.temporal_mask)})
        out[name + "_cell_scopes"] = np.array(sorted({c[0] for c in shim.GRUCell.CALLS}))

    recurrent_case("sentence", [5], [(4, "bidirectional")], False, False, True, False)
    recurrent_case("deep", [3, 1], [(4, "forward"), (4, "backward"), (2, "bidirectional"), (3, "bidirectional")],
                   True, True, True, True)
    recurrent_case("plain", [4], [(3, "backward"), (3, "forward")], True, False, False, False)
    recurrent_case("nematus", [5], [(4, "bidirectional", "NematusGRU"), (3, "forward", "NematusGRU")],
                   False, False, True, False)
    # the encoder of the reference's tests/nematus.ini: an LSTM layer under two Nematus GRU layers
    recurrent_case("mixed", [5], [(4, "forward", "LSTM"), (4, "backward", "NematusGRU"),
                                  (3, "bidirectional", "NematusGRU"), (2, "bidirectional", "LSTM")],
                   True, True, True, False)

    # ---- multi-source Transformer decoder layers: the four encoder-attention combination strategies
    #      (attention/transformer_cross_layer.py:12-263) over two encoders; the variables are drawn on demand
    #      and recorded, so the names are the reference's -----------------------------------------------------------
    enc_a, mask_a = f32(3, 5, dim), np.array([[1, 1, 1, 1, 1], [1, 1, 1, 0, 0], [1, 0, 0, 0, 0]], np.float32)
    enc_b, mask_b = f32(3, 4, dim), np.array([[1, 1, 0, 0], [1, 1, 1, 1], [1, 1, 1, 0]], np.float32)
    ms_in = f32(3, 6, dim)
    ms_mask = np.array([[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 0, 0], [1, 1, 0, 0, 0, 0]], np.float32)
    out.update({"ms_enc_a": enc_a, "ms_mask_a": mask_a, "ms_enc_b": enc_b, "ms_mask_b": mask_b, "ms_in": ms_in,
                "ms_mask": ms_mask})
    shim.AUTO[0] = np.random.RandomState(23)
    for strategy in ("serial", "parallel", "flat", "hierarchical"):
        name = "tms_" + strategy
        first = len(shim.VARIABLES)
        ms = object.__new__(TransformerDecoder)
        ms.__dict__.update(dict(
            encoders=[None, None], ff_hidden_size=ff, n_heads_self=heads, depth=2,
            n_heads_enc=[3, 3] if strategy == "flat" else [3, 2], attention_dropout_keep_prob=[1.0, 1.0],
            self_att_dropout_keep_prob=1.0, dropout_keep_prob=1.0, use_att_transform_bias=False,
            attention_combination_strategy=strategy, n_heads_hier=4 if strategy == "hierarchical" else None,
            encoder_states=lambda: [shim.t(enc_a), shim.t(enc_b)],
            encoder_masks=lambda: [shim.t(mask_a), shim.t(mask_b)],
            _embedding_size=dim, embeddings_source=None, train_mode=None,
            _variable_scope=shim.VarScope(name), _reuse=None, _name=name))
        with ms.use_scope():
            last = ms.layer(2, shim.t(ms_in), shim.t(ms_mask))
        out[name + "_states"] = np.asarray(last.temporal_states)
        for vname in list(shim.VARIABLES)[first:]:
            out["mv::" + vname] = shim.VARIABLES[vname]
    shim.AUTO[0] = None

    # ---- TransformerEncoder with its options: target-space embedding, projection biases, no position signal,
    #      cross-attention to another encoder (encoders/transformer.py:170-330); variables drawn on demand --------
    tx_in = f32(3, 6, dim)
    tx_mask = np.array([[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 0, 0], [1, 1, 0, 0, 0, 0]], np.float32)
    tx_other, tx_other_mask = f32(3, 4, dim), np.array([[1, 1, 1, 0], [1, 1, 1, 1], [1, 0, 0, 0]], np.float32)
    out.update({"tx_in": tx_in, "tx_mask": tx_mask, "tx_other": tx_other, "tx_other_mask": tx_other_mask})
    from neuralmonkey.model.stateful import TemporalStateful

    class _OtherEncoder(TemporalStateful):
        temporal_states = property(lambda self: shim.t(tx_other))
        temporal_mask = property(lambda self: shim.t(tx_other_mask))

    shim.AUTO[0] = np.random.RandomState(29)
    tf_mod = sys.modules["tensorflow"]
    tf_mod.gather = lambda params, indices: shim.t(np.asarray(params)[int(np.asarray(indices))])
    tf_mod.variance_scaling_initializer = lambda *a, **k: None
    for case, opts in (("space", dict(target_space_id=5, use_att_transform_bias=True, use_positional_encoding=True,
                                      input_for_cross_attention=None, n_cross_att_heads=None)),
                       ("cross", dict(target_space_id=None, use_att_transform_bias=False, use_positional_encoding=False,
                                      input_for_cross_attention=_OtherEncoder(), n_cross_att_heads=2))):
        name = "tx_" + case
        first = len(shim.VARIABLES)
        enc2 = object.__new__(TransformerEncoder)
        enc2.__dict__.update(dict(
            input_sequence=types.SimpleNamespace(temporal_states=shim.t(tx_in), temporal_mask=shim.t(tx_mask), dimension=dim),
            ff_hidden_size=ff, depth=2, n_heads=heads, dropout_keep_prob=1.0, attention_dropout_keep_prob=1.0,
            train_mode=None, _variable_scope=shim.VarScope(name), _reuse=None, _name=name, **opts))
        out[name + "_states"] = np.asarray(enc2.temporal_states)
        out[name + "_output"] = np.asarray(enc2.output)
        for vname in list(shim.VARIABLES)[first:]:
            out["xv::" + vname] = shim.VARIABLES[vname]
    shim.AUTO[0] = None

    # ---- WHERE dropout is applied: the same parts in training mode with keep_prob 0.5 and a deterministic
    #      mask (tf_numpy_shim.feature_dropout_mask) in place of tf.nn.dropout -----------------------------------
    shim.AUTO[0] = np.random.RandomState(31)
    train_flag = shim.t(np.bool_(True))
    d_ids = np.array([[4, 5, 6, 7, 8, 3], [3, 4, 5, 0, 0, 0], [8, 7, 0, 0, 0, 0]], np.int64)
    d_seq = object.__new__(EmbeddedFactorSequence)
    d_seq.__dict__.update(dict(
        _variable_scope=shim.VarScope("dr_enc_input"), _reuse=None, _name="dr_enc_input",
        vocabularies=[list(range(9))], vocabulary_sizes=[9], data_ids=["f0"], embedding_sizes=[6],
        scale_embeddings_by_depth=False, embeddings_source=None, trainable=True,
        _input_factor_indices_cached_placeholder=[shim.t(d_ids)]))
    first = len(shim.VARIABLES)
    d_enc = object.__new__(RecurrentEncoder)
    d_enc.__dict__.update(dict(
        _variable_scope=shim.VarScope("dr_enc"), _reuse=None, _name="dr_enc", input_sequence=d_seq,
        dropout_keep_prob=0.5, train_mode=train_flag, rnn_specs=[_make_rnn_spec(4, "bidirectional"), _make_rnn_spec(8, "forward")],
        add_residual=True, add_layer_norm=False, include_final_layer_norm=True))
    out.update({"dr_ids": d_ids, "dr_enc_states": np.asarray(d_enc.temporal_states), "dr_enc_output": np.asarray(d_enc.output)})
    d_att = object.__new__(Attention)
    d_att.__dict__.update(dict(
        _variable_scope=shim.VarScope("dr_att"), _reuse=None, _name="dr_att", _state_size=7, batch_size=3, _histories={},
        encoder=d_enc, dropout_keep_prob=0.5, train_mode=train_flag))
    d_gold = np.array([[5, 6, 7, 2, 0], [7, 8, 9, 4, 2], [4, 2, 0, 0, 0]], np.int64).T
    d_dec = object.__new__(Decoder)
    d_dec.__dict__.update(dict(
        vocabulary=list(range(12)), supress_unk=False, max_output_len=5, batch_size=3, label_smoothing=None,
        dropout_keep_prob=0.5, train_mode=train_flag, _embedding_size=5, embeddings_source=None, tie_embeddings=False,
        _variable_scope=shim.VarScope("dr_dec"), _reuse=None, _name="dr_dec",
        _go_symbols_cached_placeholder=shim.t(np.full((3,), 1, np.int64)),
        _train_inputs_cached_placeholder=shim.t(d_gold), encoders=[d_enc],
        _output_projection_spec=maxout_output(5, 0.5), _conditional_gru=False, _attention_on_input=False,
        _rnn_cell_str="GRU", _rnn_size=6, _encoder_projection=None, attentions=[d_att],
        step_scope=shim.VarScope("dr_dec/attention_decoder"), encoder_states=lambda: [], encoder_masks=lambda: [],
        input_projection=lambda *args: LoopState(*args).feedables.embedded_input))
    out["dr_gold"] = d_gold
    out["dr_dec_initial_state"] = np.asarray(d_dec.initial_state)
    out["dr_dec_train_logits"] = np.asarray(d_dec.train_logits)
    out["dr_dec_train_loss"] = np.asarray(d_dec.train_loss)
    for vname in list(shim.VARIABLES)[first:]:
        out["dv::" + vname] = shim.VARIABLES[vname]
    # the Transformer encoder and the decoder's training pass with dropout 0.5 everywhere, attention weights
    # included (their callback is the same nn.utils.dropout)
    first = len(shim.VARIABLES)
    t_ids = np.array([[4, 5, 6, 7, 8], [3, 4, 5, 0, 0], [8, 0, 0, 0, 0]], np.int64)
    t_seq = object.__new__(EmbeddedFactorSequence)
    t_seq.__dict__.update(dict(
        _variable_scope=shim.VarScope("dt_input"), _reuse=None, _name="dt_input",
        vocabularies=[list(range(9))], vocabulary_sizes=[9], data_ids=["f0"], embedding_sizes=[dim],
        scale_embeddings_by_depth=True, embeddings_source=None, trainable=True, dimension=dim,
        _input_factor_indices_cached_placeholder=[shim.t(t_ids)]))
    t_enc = object.__new__(TransformerEncoder)
    t_enc.__dict__.update(dict(
        input_sequence=t_seq, ff_hidden_size=ff, depth=2, n_heads=heads, dropout_keep_prob=0.5,
        attention_dropout_keep_prob=0.5, target_space_id=None, use_att_transform_bias=False,
        use_positional_encoding=True, input_for_cross_attention=None, n_cross_att_heads=None, train_mode=train_flag,
        _variable_scope=shim.VarScope("dt_enc"), _reuse=None, _name="dt_enc"))
    out.update({"dt_ids": t_ids, "dt_enc_states": np.asarray(t_enc.temporal_states)})
    t_gold = np.array([[5, 6, 2, 0], [7, 8, 9, 2], [4, 2, 0, 0]], np.int64).T
    t_dec = object.__new__(TransformerDecoder)
    t_dec.__dict__.update(dict(
        encoders=[t_enc], ff_hidden_size=ff, n_heads_self=heads, n_heads_enc=[2], depth=2,
        attention_dropout_keep_prob=[0.5], self_att_dropout_keep_prob=0.5, dropout_keep_prob=0.5,
        use_att_transform_bias=False, attention_combination_strategy="serial", n_heads_hier=None,
        encoder_states=lambda: [t_enc.temporal_states], encoder_masks=lambda: [t_enc.temporal_mask],
        _embedding_size=dim, embeddings_source=None, train_mode=train_flag, tie_embeddings=True, label_smoothing=None,
        vocabulary=list(range(11)), supress_unk=False, max_output_len=4, batch_size=3,
        _go_symbols_cached_placeholder=shim.t(np.full((3,), 1, np.int64)),
        _train_inputs_cached_placeholder=shim.t(t_gold),
        _variable_scope=shim.VarScope("dt_dec"), _reuse=None, _name="dt_dec"))
    out["dt_gold"] = t_gold
    out["dt_dec_train_logits"] = np.asarray(t_dec.train_logits)
    out["dt_dec_train_loss"] = np.asarray(t_dec.train_loss)
    for vname in list(shim.VARIABLES)[first:]:
        out["dw::" + vname] = shim.VARIABLES[vname]
    shim.AUTO[0] = None

    # ---- the trainer's host logic: GenericTrainer.regularization_losses / differentiable_loss_sum /
    #      gradients (per-tensor clip_by_norm) / collect_results (trainers/generic_trainer.py:84-195,27-50),
    #      around an optimizer stand-in that hands back given gradients -------------------------------------------
    from neuralmonkey.trainers.generic_trainer import GenericTrainer
    names = ["enc/rnn/gates/kernel:0", "enc/rnn/gates/bias:0", "dec/attn_projection_bias:0", "dec/attn_bias:0",
             "dec/state_to_word_W:0", "dec/state_to_word_b:0", "vgg_16/conv1/conv1_1/weights:0",
             "enc/LayerNorm/gamma:0", "dec/Bias_like/kernel:0", "resnet_thing/w:0", "Inception/w:0"]
    shim.TRAINABLE[:] = []
    grads = []
    for i, name in enumerate(names):
        var = shim.t(f32(3, 4 + i, scale=0.7))
        var.name = name
        shim.TRAINABLE.append(var)
        grads.append(None if i == 3 else shim.t(f32(3, 4 + i, scale=[0.05, 2.0][i % 2])))
        out["trn_var::" + name] = np.asarray(var)
        if grads[-1] is not None:
            out["trn_grad::" + name] = np.asarray(grads[-1])
    seen = {}

    class FakeOptimizer:
        def compute_gradients(self, loss, var_list):
            seen["loss"], seen["vars"] = loss, [v.name for v in var_list]
            return list(zip(grads, var_list))

    objectives = [types.SimpleNamespace(name="dec_a", loss=shim.t(np.float32(2.5)), gradients=None, weight=None),
                  types.SimpleNamespace(name="dec_b", loss=shim.t(np.float32(1.25)), gradients=None, weight=0.3)]
    trainer = object.__new__(GenericTrainer)
    trainer.__dict__.update(dict(objectives=objectives, l1_weight=1e-4, l2_weight=1e-8, clip_norm=1.0, var_scopes=None,
                                 var_collection="trainable_variables", optimizer=FakeOptimizer()))
    l1, l2 = trainer.regularization_losses
    clipped = trainer.gradients
    out.update({"trn_l1": np.float32(l1), "trn_l2": np.float32(l2), "trn_diff_loss": np.float32(seen["loss"]),
                "trn_objective_values": np.asarray([np.float32(v) for v in trainer.objective_values]),
                "trn_var_list": np.array(seen["vars"]), "trn_clipped_names": np.array([v.name for _, v in clipped])})
    for grad, var in clipped:
        out["trn_clipped::" + var.name] = np.asarray(grad)
    ex = object.__new__(GenericTrainer.Executable)
    ex._executor, ex.summaries = trainer, False
    ex.collect_results([{"losses": [2.5, 1.25, float(l1), float(l2)], "batch_size": 7}])
    out["trn_loss_names"] = np.array(list(ex.result.losses))
    trainer2 = object.__new__(GenericTrainer)
    trainer2.__dict__.update(dict(var_scopes=["enc", "dec/state"], var_collection="trainable_variables"))
    out["trn_scoped_var_list"] = np.array([v.name for v in trainer2.var_list])

    # ---- the reference's RNN Decoder with scaled-dot attention OBJECTS (attention/scaled_dot_product.py:246-402;
    #      tests/post-edit.ini): a MultiHeadAttention whose keys and values come from different encoders plus a
    #      ScaledDotProdAttention, both queried by one GRU decoder.  Records which dense layers the run created
    #      (the head projections land in the DECODER's step scope) and the per-head histories.  Appended after
    #      every other case so that their random draws stay what they were. ---------------------------------------
    from neuralmonkey.attention.scaled_dot_product import MultiHeadAttention, ScaledDotProdAttention

    def rnn_multihead_case(tag, heads):
        nb, tx, hsz, vsz, max_len = 3, 5, 12, 12, 5
        esz = hsz                                         # default tanh projection: output dimension = rnn_size
        dname = "md_" + tag
        first = len(shim.VARIABLES)
        var = shim.VARIABLES
        var[dname + "/initial_state/encoders_projection/kernel"] = f32(2 * hsz, hsz, scale=0.4)
        var[dname + "/initial_state/encoders_projection/bias"] = f32(hsz, scale=0.3)
        cell = dname + "/attention_decoder/OrthoGRUCell/"
        var[cell + "gates/kernel"], var[cell + "gates/bias"] = f32(esz + hsz, 2 * hsz, scale=0.5), 1.0 + f32(2 * hsz, scale=0.2)
        var[cell + "candidate/kernel"], var[cell + "candidate/bias"] = f32(esz + hsz, hsz, scale=0.5), f32(hsz, scale=0.2)
        proj = dname + "/attention_decoder/dense/"
        var[proj + "kernel"], var[proj + "bias"] = f32(hsz + esz + 2 * hsz, hsz, scale=0.4), f32(hsz, scale=0.3)
        if heads > 1:
            for name in ("query_proj", "keys_proj", "vals_proj", "output_proj"):
                var["{}/attention_decoder/{}/kernel".format(dname, name)] = f32(hsz, hsz, scale=0.4)
        for name in list(var)[first:]:
            out["mv::" + name] = var[name]
        keys, values = f32(nb, tx, hsz), f32(nb, tx, hsz)
        enc_outs = [f32(nb, hsz), f32(nb, hsz)]
        amask = np.array([[1, 1, 1, 1, 0], [1, 1, 1, 1, 1], [1, 1, 0, 0, 0]], np.float32)
        dec_w, dec_b, table = f32(hsz, vsz, scale=0.8), f32(vsz, scale=0.3), f32(vsz, esz)
        gold = np.array([[5, 6, 7, 2, 0], [7, 8, 9, 4, 2], [4, 2, 0, 0, 0]], np.int64).T
        pre = "md_{}_".format(tag)
        out.update({pre + "keys": keys, pre + "values": values, pre + "mask": amask, pre + "w": dec_w, pre + "b": dec_b,
                    pre + "table": table, pre + "gold": gold, pre + "enc_out0": enc_outs[0], pre + "enc_out1": enc_outs[1]})
        mha = object.__new__(MultiHeadAttention)
        mha.__dict__.update(dict(
            _variable_scope=shim.VarScope("ma_" + tag), _reuse=None, _name="ma_" + tag, n_heads=heads,
            dropout_keep_prob=1.0, train_mode=None, batch_size=nb, _histories={},
            _attention_keys_cached_placeholder=shim.t(keys), _attention_values_cached_placeholder=shim.t(values),
            _attention_mask_cached_placeholder=shim.t(amask)))
        sdp_att = object.__new__(ScaledDotProdAttention)
        sdp_att.__dict__.update(dict(
            _variable_scope=shim.VarScope("sa_" + tag), _reuse=None, _name="sa_" + tag, n_heads=1,
            dropout_keep_prob=1.0, train_mode=None, batch_size=nb, _histories={},
            _attention_keys_cached_placeholder=shim.t(keys), _attention_values_cached_placeholder=shim.t(keys),
            _attention_mask_cached_placeholder=shim.t(amask)))
        rd = object.__new__(Decoder)
        rd.__dict__.update(dict(
            vocabulary=list(range(vsz)), supress_unk=False, max_output_len=max_len, batch_size=nb, label_smoothing=None,
            dropout_keep_prob=1.0, train_mode=None, _embedding_size=esz, embeddings_source=None,
            _variable_scope=shim.VarScope(dname), _reuse=None, _name=dname,
            _decoding_w_cached_placeholder=shim.t(dec_w), _decoding_b_cached_placeholder=shim.t(dec_b),
            _embedding_matrix_cached_placeholder=shim.t(table),
            _go_symbols_cached_placeholder=shim.t(np.full((nb,), 1, np.int64)),
            _train_inputs_cached_placeholder=shim.t(gold),
            encoders=[types.SimpleNamespace(output=shim.t(e)) for e in enc_outs],
            _output_projection_spec=None, _conditional_gru=False, _attention_on_input=False, _rnn_cell_str="GRU",
            _rnn_size=hsz, _encoder_projection=None, attentions=[mha, sdp_att],
            step_scope=shim.VarScope(dname + "/attention_decoder"), encoder_states=lambda: [], encoder_masks=lambda: [],
            input_projection=lambda *args: LoopState(*args).feedables.embedded_input))
        shim.USED[:] = []
        out[pre + "context_sizes"] = np.array([mha.context_vector_size, sdp_att.context_vector_size])
        out[pre + "train_logits"] = np.asarray(rd.train_logits)
        out[pre + "train_loss"] = np.asarray(rd.train_loss)
        out[pre + "train_rnn_outputs"] = np.asarray(rd.train_loop_result.histories.other.rnn_outputs)
        out[pre + "run_logits"] = np.asarray(rd.runtime_logits)
        out[pre + "run_symbols"] = np.asarray(rd.runtime_loop_result.histories.output_symbols)
        for mode in ("train", "run"):
            for i in range(heads):
                out["{}{}_mha_head{}".format(pre, mode, i)] = np.asarray(mha.histories["{}_{}_head{}".format(dname, mode, i)])
            out["{}{}_sdp_head0".format(pre, mode)] = np.asarray(sdp_att.histories["{}_{}_head0".format(dname, mode)])
        out[pre + "history_keys"] = np.array(sorted(mha.histories) + sorted(sdp_att.histories))
        out[pre + "dense_names"] = np.array(sorted(set(shim.USED)))

    rnn_multihead_case("h3", 3)
    rnn_multihead_case("h1", 1)
    np.savez_compressed(os.path.join(HERE, "tf_shim_golden.npz"), **out)
    print(sorted(out))


if __name__ == "__main__":
    main()
