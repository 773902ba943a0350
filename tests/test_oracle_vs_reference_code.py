"""The oracle against outputs of the reference's OWN functions (tests/golden/tf_shim_golden.npz:
neuralmonkey/{encoders/transformer,attention/scaled_dot_product,tf_utils,nn/projection,functions,
decoders/beam_search_decoder}.py executed over a numpy stand-in for the TensorFlow ops they call -
tests/golden/tf_numpy_shim.py, make_tf_shim_golden.py).  This pins the oracle's restatement of the
op order, constants, masking rules and reshapes to the reference's code itself."""
import os

import numpy as np
import pytest
import torch

from oracle import nm_oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_shim_golden.npz"))


def _t(name):
    return torch.from_numpy(G[name])


def test_position_signal():
    for dim, length in ((6, 7), (512, 50), (9, 4)):
        want = G["pos_{}_{}".format(dim, length)]
        got = O.position_signal(dim, length).numpy()
        # fp32 exp / sin of arguments up to ~50: libm differences of an ulp in the argument show up as ~1e-5
        assert got.shape == want.shape and np.abs(got - want).max() < 2e-5


def test_attention_masks_and_single_head():
    q, k, v, mask = _t("att_q"), _t("att_k"), _t("att_v"), _t("att_mask")
    ctx, weights = O.multihead_attention({}, "", q, k, v, mask, heads=1)
    assert np.abs(ctx.numpy() - G["att1_ctx"]).max() < 2e-6
    assert np.abs(weights.numpy() - G["att1_w"]).max() < 2e-6
    # padded keys get probability exactly zero, as in the reference's e*m + (1-m)*(-1e9)
    assert float(weights[1, 0, :, 3:].abs().max()) == 0.0 and float(G["att1_w"][1, 0, :, 3:].max()) == 0.0
    # the two masking helpers, applied to given energies
    e = _t("energies")
    m4 = mask.unsqueeze(1).unsqueeze(1)
    assert np.array_equal((e * m4 + (1.0 - m4) * -1e9).numpy(), G["mask_energies"])
    sq = _t("energies_sq")
    keep = torch.tril(torch.ones(5, 5, dtype=torch.bool))
    assert np.array_equal(torch.where(keep, sq, torch.full_like(sq, -1e9)).numpy(), G["mask_future"])
    assert np.array_equal(O._split_for_heads(q, 3, 4).numpy(), G["split_heads"])


def test_multi_head_cross_and_masked_self_attention():
    p = {"s/{}/kernel".format(n): _t("dense_" + n) for n in ("query_proj", "keys_proj", "vals_proj", "output_proj")}
    ctx, weights = O.multihead_attention(p, "s", _t("att_q"), _t("att_k"), _t("att_v"), _t("att_mask"), heads=3)
    assert np.abs(ctx.numpy() - G["att3_ctx"]).max() < 5e-6
    assert np.abs(weights.numpy() - G["att3_w"]).max() < 2e-6
    qs = _t("self_q")
    ctx, weights = O.multihead_attention(p, "s", qs, qs, qs, _t("self_mask"), heads=3, masked=True)
    assert np.abs(ctx.numpy() - G["self_ctx"]).max() < 5e-6
    assert np.abs(weights.numpy() - G["self_w"]).max() < 2e-6
    # dropout on the attention weights (:208-214): after the softmax, before the values; the reference
    # returns the DROPPED weights
    ctx, weights = O.multihead_attention(p, "s", qs, qs, qs, _t("self_mask"), heads=3, masked=True,
                                         drop_mask=_t("drop_mask"))
    assert np.abs(ctx.numpy() - G["drop_ctx"]).max() < 5e-6
    assert np.abs(weights.numpy() - G["drop_w"]).max() < 2e-6


def test_layer_norm_and_maxout():
    got = O.layer_norm(_t("ln_x"), _t("ln_gamma"), _t("ln_beta")).numpy()
    assert np.abs(got - G["ln_y"]).max() < 5e-6
    spec = O.RNNDecoderSpec("d", "a", 5, "maxout", False)
    pre = "d/attention_decoder/MaxoutProjection/MaxoutProjection/"
    p = {pre + "kernel": _t("maxout_kernel"), pre + "bias": _t("maxout_bias")}
    x = _t("maxout_in")
    got = O.output_projection(p, spec, x[:, :3], x[:, 3:5], x[:, 5:]).numpy()   # concat of the three = x
    assert np.abs(got - G["maxout_out"]).max() < 2e-6


def test_beam_reordering_helpers():
    """gather_flat (the per-hypothesis re-ordering of every decoder feedable), partial_transpose and
    append_tensor as the oracle's beam search uses them."""
    state, beams = _t("gf_state"), torch.from_numpy(G["gf_beam_ids"]).long()
    batch, beam = beams.shape
    flat = (torch.arange(batch).unsqueeze(1) * beam + beams).reshape(-1)       # oracle: beam_search()
    assert np.array_equal(state[flat].numpy(), G["gf_out"])
    assert np.array_equal(_t("pt_in").transpose(0, 1).numpy(), G["pt_out"])
    hist = _t("pt_in")
    assert np.array_equal(torch.cat([hist, (hist[0] * 2).unsqueeze(0)], 0).numpy(), G["append_out"])


def test_length_penalty_and_noam_schedule():
    from neuralmonkey_b200.functions import noam_decay
    lengths = torch.from_numpy(G["lp_lengths"])
    for alpha in (0.0, 0.6, 1.0):
        got = O.length_penalty(lengths, alpha).numpy()
        want = G["lp_{}".format(alpha)]
        assert np.abs(got - want).max() <= 2.4e-7 * np.abs(want).max()        # within 1 ulp of fp32 pow
    sched = noam_decay(0.2, 6, 111)
    for step, want in zip(G["noam_steps"].tolist(), G["noam_values"].tolist()):
        assert sched(step) == pytest.approx(want, rel=1e-6, abs=1e-12)


def test_bahdanau_attention_step():
    """Attention.attention (attention/feed_forward.py:125-166): query projection + bias, tanh energies,
    softmax over ALL positions, then mask and renormalise with +1e-8; context over the states."""
    p = {"a/Attention/attn_query_projection": _t("bah_attn_query_projection"),
         "a/attn_key_projection": _t("bah_attn_key_projection"), "a/attn_similarity_v": _t("bah_attn_similarity_v"),
         "a/attn_projection_bias": _t("bah_attn_projection_bias"), "a/attn_bias": _t("bah_attn_bias")}
    states, query = _t("bah_states"), _t("bah_query")
    hidden = O.bahdanau_precompute(p, "a", states)
    for label, mask in (("masked", _t("bah_mask")), ("nomask", None)):
        ctx, weights = O.bahdanau_step(p, "a", query, hidden, states, mask)
        assert np.abs(weights.numpy() - G["bah_w_" + label]).max() < 2e-6, label
        assert np.abs(ctx.numpy() - G["bah_ctx_" + label]).max() < 5e-6, label
    # a sentence of length 1 puts (almost) all weight on its only position: 1 / (1 + 1e-8 / w)
    assert abs(float(G["bah_w_masked"][2, 0]) - 1.0) < 1e-6 and float(np.abs(G["bah_w_masked"][2, 1:]).max()) == 0.0


def test_decoder_projections():
    """linear_encoder_projection (encoder outputs concatenated in list order -> dense) and the two
    output projections (concat of [cell output, embedded input, contexts] -> dense+tanh / maxout)."""
    spec_t = O.RNNDecoderSpec("d", "a", 5, "tanh", False)
    spec_m = O.RNNDecoderSpec("d", "a", 5, "maxout", False)
    p = {"d/initial_state/encoders_projection/kernel": _t("proj_kernel"),
         "d/initial_state/encoders_projection/bias": _t("proj_bias"),
         "d/attention_decoder/dense/kernel": _t("op_tanh_kernel"), "d/attention_decoder/dense/bias": _t("op_tanh_bias"),
         "d/attention_decoder/MaxoutProjection/MaxoutProjection/kernel": _t("op_max_kernel"),
         "d/attention_decoder/MaxoutProjection/MaxoutProjection/bias": _t("op_max_bias")}
    init = O.decoder_initial_state(p, spec_t, torch.cat([_t("proj_enc_a"), _t("proj_enc_b")], 1))
    assert np.abs(init.numpy() - G["proj_init"]).max() < 2e-6
    cell, emb, ctx = _t("op_cell"), _t("op_emb"), _t("op_ctx")
    assert np.abs(O.output_projection(p, spec_t, cell, emb, ctx).numpy() - G["op_tanh_out"]).max() < 2e-6
    assert np.abs(O.output_projection(p, spec_m, cell, emb, ctx).numpy() - G["op_max_out"]).max() < 2e-6


def test_transformer_encoder_and_decoder_stacks():
    """TransformerEncoder.temporal_states / output and TransformerDecoder.layer(depth, ...) - the
    reference's whole layer stacks (encoders/transformer.py:198-318, decoders/transformer.py:270-387),
    every variable fetched by its full TF scope name: the names ARE the arena's variable names."""
    p = {k[4:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("tv::")}
    enc = O.transformer_encoder(p, "tenc", _t("tenc_in"), _t("tenc_mask"), depth=2, heads=3)
    assert np.abs(enc["states"].numpy() - G["tenc_states"]).max() < 2e-5
    assert np.abs(enc["output"].numpy() - G["tenc_output"]).max() < 5e-5       # unmasked SUM over time
    spec = O.TransformerDecoderSpec("tdec", 2, 3, 2, 9)
    states = O.transformer_decoder_stack(p, spec, _t("tdec_in"), _t("tdec_mask"), _t("tenc_states"), _t("tenc_mask"))
    assert np.abs(states.numpy() - G["tdec_states"]).max() < 2e-5
    # every dense kernel the reference looked up exists under exactly that name, and none is unused
    used = set(G["transformer_dense_names"].tolist())
    kernels = {n for n in p if n.endswith("/kernel")}
    assert used == kernels


def test_product_variable_names_follow_the_reference_scopes():
    """The names the reference's code asked the (stand-in) variable store for are the names this
    package declares for the same model parts (checkpoint / importer compatibility)."""
    import re
    names = {k[4:] for k in G.files if k.startswith("tv::")}
    enc_names = {re.sub(r"^tenc/", "", n) for n in names if n.startswith("tenc/")}
    dec_names = {re.sub(r"^tdec/", "", n) for n in names if n.startswith("tdec/")}
    from neuralmonkey_b200.params import ParameterArena
    from neuralmonkey_b200 import runtime
    from neuralmonkey_b200.attention.scaled_dot_product import declare_attention
    from neuralmonkey_b200.attention.transformer_cross_layer import declare_cross
    from neuralmonkey_b200.encoders.transformer import declare_feedforward, declare_layer_norm

    class Part:                      # collects local names the way Parameterized.declare does
        def __init__(self):
            self.names = set()

        def declare(self, local_name, shape, initializer=None, trainable=True, absolute=False):
            self.names.add(local_name)

    enc_part, dec_part = Part(), Part()
    for i in range(2):
        scope = "layer_{}".format(i)
        declare_layer_norm(enc_part, scope + "/self_attention", 12)
        declare_attention(enc_part, scope + "/self_attention", 12, 12, 3)
        declare_feedforward(enc_part, scope + "/feedforward", 12, 20)
        declare_layer_norm(dec_part, scope + "/self_attention", 12)
        declare_attention(dec_part, scope + "/self_attention", 12, 12, 3)
        declare_cross(dec_part, scope + "/encdec_attention", "serial", 12, [2])
        declare_feedforward(dec_part, scope + "/feedforward", 12, 20)
    for part in (enc_part, dec_part):
        part.names |= {"LayerNorm/gamma", "LayerNorm/beta"}
    assert enc_part.names == enc_names
    assert dec_part.names == dec_names


@pytest.mark.parametrize("case", ["a", "b"])
def test_one_beam_search_step(case):
    """BeamSearchDecoder.get_body()() - the reference's beam step (finished-row masking, hypothesis
    scores with the length penalty, top-k over beam*vocabulary, gathers of lengths / UN-normalised
    logprob sums / finished flags / every decoder feedable, token history re-ordering, embedding of
    the chosen words, log-softmax of the parent's next logits) against the oracle.  Case b holds
    exact score ties inside a sentence."""
    pre = "beam_{}_".format(case)
    alpha = float(G[pre + "alpha"])
    scores, words, beams, lsum, lens, fin = O.beam_step(_t(pre + "prev_logprobs"), _t(pre + "logprob_sum"),
                                                        _t(pre + "lengths"), _t(pre + "finished"), alpha)
    assert np.array_equal(lens.numpy(), G[pre + "out_lengths"])
    assert np.array_equal(fin.numpy(), G[pre + "out_finished"])
    assert np.abs(scores.numpy() - G[pre + "out_scores"]).max() < 2e-6
    assert np.abs(lsum.numpy() - G[pre + "out_logprob_sum"]).max() < 2e-6
    bsz, beam = beams.shape
    # token history: re-ordered by the surviving beams, the new words appended (oracle: beam_search)
    tokens = _t(pre + "token_ids")
    tokens = tokens[:, torch.arange(bsz).unsqueeze(1), beams.long()]
    tokens = torch.cat([tokens, words.unsqueeze(0)], 0)
    assert np.array_equal(tokens.numpy(), G[pre + "out_token_ids"])
    flat = (torch.arange(bsz).unsqueeze(1) * beam + beams.long()).reshape(-1)
    assert np.array_equal(_t(pre + "state_feed")[flat].numpy(), G[pre + "out_state_feed"])
    assert np.array_equal(_t(pre + "emb_table")[words.reshape(-1)].numpy(), G[pre + "out_embedded"])
    assert np.array_equal(fin.reshape(-1).numpy(), G[pre + "out_dec_finished"])
    want_lp = G[pre + "out_prev_logprobs"]
    got_lp = torch.log_softmax(_t(pre + "next_logits"), -1).reshape(want_lp.shape).numpy()
    assert np.abs(got_lp - want_lp).max() < 2e-6


@pytest.mark.parametrize("mode", ["train", "run", "run_unk", "run_eos"])
def test_shared_decoding_loop(mode):
    """AutoregressiveDecoder.get_initial_loop_state / loop_continue_criterion / get_body
    (autoregressive.py:381-519) driven step by step around a stand-in next_state: logits with the
    -1e9 <unk> column, gold vs argmax feedback over the full vocabulary, `symbol *= unfinished`,
    `finished |= symbol == </s>`, mask = not finished AFTER the step, stop when all finished or at
    max_output_len - against the loop the oracle's three decoders share."""
    states = _t("loop_states")
    w, b, table = _t("loop_w"), _t("loop_" + mode + "_bias"), _t("loop_table")
    supress = mode != "run"
    gold = torch.from_numpy(G["loop_gold"]) if mode == "train" else None
    counter = {"step": 0}

    def next_output(_embedded, _finished):
        out = states[counter["step"]]
        counter["step"] += 1
        return out, None

    def to_logits(out):
        logits = out @ w + b
        if supress:
            pen = torch.zeros(logits.shape[-1])
            pen[O.UNK] = -1e9
            logits = logits + pen
        return logits

    seen_inputs = []

    def embed(ids):
        seen_inputs.append(ids.clone())
        return table[ids]

    hist = O.autoregressive_loop(next_output, to_logits, embed, states.shape[1], states.shape[0], gold)
    steps = int(G["loop_{}_steps".format(mode)])
    assert len(hist["symbols"]) == steps
    assert np.array_equal(torch.stack(hist["symbols"]).numpy(), G["loop_{}_symbols".format(mode)])
    assert np.array_equal(torch.stack(hist["mask"]).numpy(), G["loop_{}_mask".format(mode)])
    assert np.abs(torch.stack(hist["logits"]).numpy() - G["loop_{}_logits".format(mode)]).max() < 2e-6
    assert np.array_equal(table[seen_inputs[-1]].numpy(), G["loop_{}_last_input".format(mode)])
    if mode == "train":
        # the product computes the symbols fed at every training step on the host, for all steps at
        # once: they are the inputs the reference's loop embedded, step by step
        from neuralmonkey_b200.decoders.autoregressive import AutoregressiveDecoder
        fed = AutoregressiveDecoder.teacher_forcing_inputs(G["loop_gold"].T.copy())       # [B, T]
        want = torch.stack(seen_inputs[:steps]).t().numpy()                                 # <s>, then fed-back gold
        assert np.array_equal(fed[:, :steps], want)


@pytest.mark.parametrize("tag,maxout,use_mask", [("maxout", True, True), ("tanh", False, False)])
def test_whole_attention_decoder(tag, maxout, use_mask):
    """decoders/decoder.py Decoder run whole (initial_state :226-251, get_initial_feedables/histories
    :360-382, next_state :279-358) under AutoregressiveDecoder's loop with the reference's own Attention,
    linear_encoder_projection and maxout / tanh output projection, in training and in greedy mode; only
    the GRU cell arithmetic inside is TensorFlow's published GRUCell restated by the shim.  Pins, against
    the oracle's decoder_train / decoder_greedy: the cell is fed (embedded_input, prev_rnn_output); the
    attention query is the cell output; the projection input order [cell, embedding, context]; the
    variable names; initial state = dense(encoder output)."""
    dname, aname = "rd_" + tag, "ra_" + tag
    p = {k[4:]: _t(k) for k in G.files if k.startswith("rv::" + dname) or k.startswith("rv::" + aname)}
    p[dname + "/word_embeddings"] = _t(dname + "_table")
    p[dname + "/state_to_word_W"], p[dname + "/state_to_word_b"] = _t(dname + "_w"), _t(dname + "_b")
    spec = O.RNNDecoderSpec(dname, aname, max_output_len=6, output_projection="maxout" if maxout else "tanh")
    enc = {"temporal_states": _t(dname + "_states"), "output": _t(dname + "_enc_out"),
           "temporal_mask": _t(dname + "_mask") if use_mask else None}
    gold = torch.from_numpy(G[dname + "_gold"])
    assert np.abs(O.decoder_initial_state(p, spec, enc["output"]).numpy()
                  - G[dname + "_train_initial_state"]).max() < 2e-6
    # the names the reference asked the variable store for
    dense = {n[:-len("/kernel")] for n in G[dname + "_dense_names"].tolist()}
    proj = "MaxoutProjection/MaxoutProjection" if maxout else "dense"
    assert dense == {dname + "/initial_state/encoders_projection", dname + "/attention_decoder/" + proj}
    assert G[dname + "_cell_scopes"].tolist() == [dname + "/attention_decoder/OrthoGRUCell/"]

    train = O.decoder_train(p, spec, enc, gold)
    key = dname + "_train_"
    assert np.abs(train["train_logits"].numpy() - G[key + "logits"]).max() < 5e-6
    assert np.abs(train["train_output_states"].numpy() - G[key + "output_states"]).max() < 5e-6
    assert np.abs(train["rnn_outputs"].numpy() - G[key + "rnn_outputs"]).max() < 5e-6
    assert np.abs(train["attention_weights"].numpy() - G[key + "att_weights"]).max() < 5e-6
    if use_mask:
        assert float(np.abs(G[key + "att_weights"][:, 3, 1:]).max()) == 0.0     # padded source positions

    run = O.decoder_greedy(p, spec, enc, gold)
    key = dname + "_run_"
    assert np.array_equal(run["output_symbols"].numpy(), G[key + "symbols"])
    assert np.array_equal(run["runtime_mask"].numpy(), G[key + "out_mask"])
    assert np.abs(run["runtime_logits"].numpy() - G[key + "logits"]).max() < 5e-6
    # the loss tensors (autoregressive.py:292-371) over tf.contrib.seq2seq.sequence_loss restated by the
    # shim: [B, T] xents weighted by the gold mask, token-mean train loss, run-time xents cut to
    # min(gold, decoded) steps and divided by the number of DECODED (not gold) unfinished positions
    assert np.abs(train["train_xents"].numpy() - G[dname + "_train_xents"]).max() < 1e-5
    assert abs(float(train["train_loss"]) - float(G[dname + "_train_loss"])) < 1e-5
    assert np.array_equal(train["train_mask"].numpy(), G[dname + "_train_mask"])
    # label smoothing: one scalar (the mean over ALL positions of the smoothed cross-entropy) times the mask
    smooth = O.decoder_train(p, spec, enc, gold, label_smoothing=0.1)
    assert np.abs(smooth["train_xents"].numpy() - G[dname + "_smooth_train_xents"]).max() < 1e-5
    assert abs(float(smooth["train_loss"]) - float(G[dname + "_smooth_train_loss"])) < 1e-5
    assert len(set(np.round(G[dname + "_smooth_train_xents"][G[dname + "_smooth_train_xents"] > 0], 4))) == 1
    assert np.abs(run["runtime_xents"].numpy() - G[dname + "_runtime_xents"]).max() < 1e-5
    assert abs(float(run["runtime_loss"]) - float(G[dname + "_runtime_loss"])) < 1e-5
    assert np.array_equal(run["decoded"].numpy(), G[dname + "_decoded"])
    assert np.abs(run["runtime_logprobs"].numpy() - G[dname + "_runtime_logprobs"]).max() < 1e-5


@pytest.mark.parametrize("tag,layers,residual,layer_norm,final_norm,scale", [
    ("sentence", [(4, "bidirectional")], False, False, True, False),
    ("deep", [(4, "forward"), (4, "backward"), (2, "bidirectional"), (3, "bidirectional")], True, True, True, True),
    ("plain", [(3, "backward"), (3, "forward")], True, False, False, False),
    ("nematus", [(4, "bidirectional", "NematusGRU"), (3, "forward", "NematusGRU")], False, False, True, False),
    ("mixed", [(4, "forward", "LSTM"), (4, "backward", "NematusGRU"), (3, "bidirectional", "NematusGRU"),
               (2, "bidirectional", "LSTM")], True, True, True, False)])
def test_recurrent_encoder(tag, layers, residual, layer_norm, final_norm, scale):
    """model/sequence.py EmbeddedFactorSequence.temporal_states / temporal_mask (:170-199) and
    encoders/recurrent.py RecurrentEncoder.rnn + rnn_layer (:71-110,180-218) run whole: factor lookup,
    sqrt(size) scaling, masking by the first factor, per-layer scopes and LayerNorm variables, which
    input the residual adds, the shared final LayerNorm, fw|bw concatenation order.  The recurrence
    underneath (dynamic_rnn, GRUCell) is TensorFlow library code restated by the shim."""
    name = "re_" + tag
    p = {k[4:]: _t(k) for k in G.files if k.startswith("ev::" + name + "/") or k.startswith("ev::" + name + "_input/")}
    factors = [torch.from_numpy(f) for f in G[name + "_ids"]]
    seq = O.embedded_sequence(p, name + "_input", factors, scale_embeddings_by_depth=scale)
    assert np.array_equal(seq["temporal_mask"].numpy(), G[name + "_mask"])
    assert np.abs(seq["temporal_states"].numpy() - G[name + "_embedded"]).max() < 1e-6
    assert np.array_equal(seq["temporal_mask"].sum(1).to(torch.int32).numpy(), G[name + "_lengths"])
    enc = O.recurrent_encoder(p, name, seq["temporal_states"], seq["temporal_mask"], layers, residual,
                              layer_norm, final_norm)
    assert np.abs(enc["temporal_states"].numpy() - G[name + "_states"]).max() < 5e-6
    assert np.abs(enc["output"].numpy() - G[name + "_output"]).max() < 5e-6
    assert np.array_equal(enc["temporal_mask"].numpy(), G[name + "_enc_mask"])
    # the scopes the reference called its cells in are the ones the oracle reads its parameters from
    cell_scopes = {k[: k.index("gates/")] for k in p if "/gates/" in k} | {
        k[: -len("kernel")] for k in p if k.endswith("lstm_cell/kernel")}
    assert set(G[name + "_cell_scopes"].tolist()) == cell_scopes


def test_product_declares_the_variables_the_reference_asks_for():
    """The names under which the product's RecurrentEncoder / Decoder / Attention declare their
    parameters are the names the reference's code looked up in the variable store (recorded by the
    shim) - what makes reference checkpoints and the oracle's parameter dictionaries interchangeable."""
    from neuralmonkey_b200.encoders.recurrent import RecurrentEncoder, _make_rnn_spec
    enc = object.__new__(RecurrentEncoder)
    want = set(G["re_deep_cell_scopes"].tolist())
    layers = [(4, "forward"), (4, "backward"), (2, "bidirectional"), (3, "bidirectional")]
    got = set()
    for i, layer in enumerate(layers):
        got.update("re_deep/" + scope + "/" for scope in RecurrentEncoder._cell_scopes(enc, i, _make_rnn_spec(*layer)))
    assert got == want


@pytest.mark.parametrize("tag,dtag,maxout,use_mask", [("rnn", "beam", True, True), ("rnn1", "beam1", False, True),
                                                      ("rnn2", "beam2", False, False)])
def test_whole_beam_search_over_the_attention_decoder(tag, dtag, maxout, use_mask):
    """BeamSearchDecoder.get_initial_loop_state / loop_continue_criterion / get_body
    (decoders/beam_search_decoder.py:218-558) run to the end around the reference's own attention
    Decoder (one sentence: the reference's RNN decoder does not tile encoder states to the beam),
    against the oracle's beam_search over its decoder_step: hypotheses, scores, lengths, finished
    flags, number of steps (max_steps reached / all finished), and the first-step log-probs."""
    dname, aname, pre = "rd_" + dtag, "ra_" + dtag, "bsearch_{}_".format(tag)
    p = {k[4:]: _t(k) for k in G.files if k.startswith("rv::" + dname + "/") or k.startswith("rv::" + aname + "/")}
    p[dname + "/word_embeddings"] = _t(dname + "_table")
    p[dname + "/state_to_word_W"], p[dname + "/state_to_word_b"] = _t(dname + "_w"), _t(dname + "_b")
    spec = O.RNNDecoderSpec(dname, aname, max_output_len=6, output_projection="maxout" if maxout else "tanh")
    states, mask = _t(dname + "_states"), (_t(dname + "_mask") if use_mask else None)
    beam, alpha, max_steps = int(G[pre + "beam"]), float(G[pre + "alpha"]), int(G[pre + "max_steps"])
    hidden = O.bahdanau_precompute(p, aname, states)
    emb = p[dname + "/word_embeddings"]

    def run_step(prev_output, words):
        output, cell, _ctx, _w = O.decoder_step(p, spec, emb[words], prev_output, hidden, states, mask)
        return cell, torch.log_softmax(O.state_to_logits(p, spec, output), -1)

    init = O.decoder_initial_state(p, spec, _t(dname + "_enc_out")).repeat_interleave(beam, 0)
    state, first = run_step(init, torch.full((beam,), O.START, dtype=torch.int64))
    assert np.abs(first.reshape(1, beam, -1).numpy() - G[pre + "init_prev_logprobs"]).max() < 5e-6
    assert np.array_equal(G[pre + "init_logprob_sum"][0, 1:], np.full(beam - 1, -1e9, np.float32))
    calls = {"n": 0}

    def step_fn(st, words, _finished):
        calls["n"] += 1
        return run_step(st, words)

    got = O.beam_search(step_fn, state, first, beam, max_steps, alpha, lambda st, idx: st[idx])
    assert calls["n"] == int(G[pre + "steps"])
    # slot 0 of the reference's token history holds the first step's greedy symbol (:301-313), which
    # the runner drops (beamsearch_runner.py:84); the hypotheses start at slot 1
    assert np.array_equal(got["token_ids"].numpy(), G[pre + "token_ids"][1:])
    assert np.array_equal(got["lengths"].numpy(), G[pre + "lengths"])
    assert np.array_equal(got["finished"].numpy(), G[pre + "finished"])
    assert np.abs(got["scores"].numpy() - G[pre + "scores"]).max() < 5e-6
    assert np.array_equal(G[pre + "init_token_ids"].reshape(-1), np.full(beam, int(first[0].argmax())))


def _transformer_setup(table_key="tloop_table"):
    p = {k[4:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("tv::")}
    p["tdec/word_embeddings"] = _t(table_key)
    spec = O.TransformerDecoderSpec("tdec", 2, 3, 2, 6, tie_embeddings=True)
    enc = {"states": _t("tenc_states"), "mask": _t("tenc_mask")}
    return p, spec, enc


def test_transformer_decoder_training_pass():
    """TransformerDecoder.train_loop_result (decoders/transformer.py:393-453) run whole: inputs are
    <s> + gold[:-1] embedded WITHOUT the position signal, the self-attention mask is the gold mask,
    logits come from the transposed embedding matrix (tie_embeddings, zero bias)."""
    p, spec, enc = _transformer_setup()
    gold = torch.from_numpy(G["tloop_gold"]).t().contiguous()                     # [B, T]
    assert np.array_equal(G["tloop_train_input_symbols"][:, 1:], gold[:, :-1].numpy())
    assert np.array_equal(G["tloop_train_input_symbols"][:, 0], np.full(gold.shape[0], O.START))
    train = O.transformer_decoder_train(p, spec, enc, gold)
    assert np.abs(train["logits"].transpose(0, 1).numpy() - G["tloop_train_logits"]).max() < 5e-5
    assert np.abs(train["states"].transpose(0, 1).numpy() - G["tloop_train_states"]).max() < 2e-5


@pytest.mark.parametrize("case,table", [("run", "tloop_table"), ("run_eos", "tloop_table_eos")])
def test_transformer_decoder_greedy_loop(case, table):
    """AutoregressiveDecoder.decoding_loop(train_mode=False) around TransformerDecoder.next_state
    (:487-516): the prefix is re-run at every step, and the self-attention mask column appended at a
    step is `not finished` as it stood BEFORE that step's symbol was chosen."""
    p, spec, enc = _transformer_setup(table)
    got = O.transformer_decoder_greedy(p, spec, enc)
    assert np.array_equal(got["symbols"].numpy(), G["tloop_{}_symbols".format(case)])
    assert np.array_equal(got["mask"].numpy(), G["tloop_{}_mask".format(case)])
    assert np.abs(got["logits"].numpy() - G["tloop_{}_logits".format(case)]).max() < 5e-5
    input_mask = G["tloop_{}_input_mask".format(case)][..., 0]
    finished_before = np.concatenate([np.zeros((1, input_mask.shape[0]), bool),
                                      ~G["tloop_{}_mask".format(case)][:-1]], 0)
    assert np.array_equal(input_mask, (~finished_before).T.astype(np.float32))


@pytest.mark.parametrize("tag,table", [("tr", "tloop_table"), ("tr_eos", "tloop_table_eos")])
def test_whole_beam_search_over_the_transformer_decoder(tag, table):
    """BeamSearchDecoder.outputs (:167-191: encoder states and masks tiled to the beam, initial loop
    state, tf.while_loop) around the reference's TransformerDecoder, three sentences."""
    p, spec, enc = _transformer_setup(table)
    pre = "bsearch_{}_".format(tag)
    beam, alpha, max_steps = int(G[pre + "beam"]), float(G[pre + "alpha"]), int(G[pre + "max_steps"])
    emb = p["tdec/word_embeddings"]
    states, emask = enc["states"].repeat_interleave(beam, 0), enc["mask"].repeat_interleave(beam, 0)
    rows = states.shape[0]

    def run(seq, mask):
        out = O.transformer_decoder_stack(p, spec, seq, mask, states, emask)
        return torch.log_softmax(O.transformer_logits(p, spec, out[:, -1]), -1)

    seq0 = emb[torch.full((rows,), O.START, dtype=torch.int64)].unsqueeze(1)
    mask0 = torch.ones(rows, 1)
    first = run(seq0, mask0)
    assert np.abs(first.reshape(-1, beam, first.shape[-1]).numpy() - G[pre + "init_prev_logprobs"]).max() < 5e-5

    def step_fn(state, words, finished):
        seq = torch.cat([state[0], emb[words].unsqueeze(1)], 1)
        mask = torch.cat([state[1], (~finished).to(emb.dtype).unsqueeze(1)], 1)
        return (seq, mask), run(seq, mask)

    got = O.beam_search(step_fn, (seq0, mask0), first, beam, max_steps, alpha,
                        lambda st, idx: (st[0][idx], st[1][idx]))
    assert got["token_ids"].shape[0] == int(G[pre + "steps"])
    assert np.array_equal(got["token_ids"].numpy(), G[pre + "token_ids"][1:])
    assert np.array_equal(got["lengths"].numpy(), G[pre + "lengths"])
    assert np.array_equal(got["finished"].numpy(), G[pre + "finished"])
    assert np.abs(got["scores"].numpy() - G[pre + "scores"]).max() < 2e-5


def test_trainer_host_logic():
    """GenericTrainer.regularization_losses / differentiable_loss_sum / gradients / collect_results
    (trainers/generic_trainer.py:27-50,84-195) around an optimizer stand-in: which variables are
    regularised ([Bb]ias anywhere in the name and vgg/Inception/resnet prefixes are not; LayerNorm
    gains and
    `state_to_word_b` are), loss = sum w_i * objective_i + l1_weight * L1 + l2_weight * L2 with weight None == 1,
    tf.clip_by_norm PER TENSOR (not global), None gradients dropped, names of the fetched losses."""
    from neuralmonkey_b200.params import is_regularizable
    variables = {k[len("trn_var::"):]: _t(k) for k in G.files if k.startswith("trn_var::")}
    l1, l2 = O.regularization(variables)
    assert abs(float(l1) - float(G["trn_l1"])) < 1e-4 * float(G["trn_l1"])
    assert abs(float(l2) - float(G["trn_l2"])) < 1e-4 * float(G["trn_l2"])
    # `state_to_word_b` IS regularised: the filter is the regex on the name, and that name has no "bias"
    want_reg = {"enc/rnn/gates/kernel:0", "dec/state_to_word_W:0", "dec/state_to_word_b:0", "enc/LayerNorm/gamma:0"}
    assert {n for n in variables if O.is_regularizable(n)} == want_reg
    assert {n for n in variables if is_regularizable(n)} == want_reg            # the product's arena flags
    assert np.array_equal(G["trn_objective_values"][:2], np.array([2.5, 1.25], np.float32))
    want_loss = 2.5 * 1.0 + 1.25 * 0.3 + 1e-4 * float(G["trn_l1"]) + 1e-8 * float(G["trn_l2"])
    assert abs(float(G["trn_diff_loss"]) - want_loss) < 1e-5
    clipped = G["trn_clipped_names"].tolist()
    assert "dec/attn_bias:0" not in clipped and len(clipped) == len(variables) - 1
    some_clipped = some_kept = False
    for name in clipped:
        grad = _t("trn_grad::" + name)
        got = O.clip_by_norm(grad, 1.0)
        assert np.abs(got.numpy() - G["trn_clipped::" + name]).max() < 1e-6
        norm = float(grad.pow(2).sum().sqrt())
        some_clipped |= norm > 1.0
        some_kept |= norm < 1.0 and np.array_equal(G["trn_clipped::" + name], grad.numpy())
    assert some_clipped and some_kept
    assert G["trn_loss_names"].tolist() == ["dec_a", "dec_b", "L1", "L2"]


@pytest.mark.parametrize("tag,cell,conditional,out_proj,enc_proj", [
    ("nematus", "NematusGRU", True, "nematus", "nematus"),
    ("cond_gru", "GRU", True, "mlp", "concat"),
    ("nematus_plain", "NematusGRU", False, "maxout", "empty"),
    ("lstm", "LSTM", False, "maxout", "linear")])
def test_decoder_variants(tag, cell, conditional, out_proj, enc_proj):
    """The decoder variants of SURVEY.md 8(f) N4, the reference's Decoder run whole with them:
    NematusGRUCell (the reference's OWN cell code: nn/ortho_gru_cell.py:57-105), the conditional GRU
    (decoder.py:303-325), nematus_output / mlp_output (output_projection.py:76-112,163-188), and the
    nematus / concat / empty initial states (encoder_projection.py:30-145)."""
    dname, aname, pre = "vd_" + tag, "va_" + tag, "vd_{}_".format(tag)
    p = {k[4:]: _t(k) for k in G.files if k.startswith("vv::" + dname + "/") or k.startswith("vv::" + aname + "/")}
    p[dname + "/word_embeddings"] = _t(pre + "table")
    p[dname + "/state_to_word_W"], p[dname + "/state_to_word_b"] = _t(pre + "w"), _t(pre + "b")
    spec = O.RNNDecoderSpec(dname, aname, max_output_len=5, output_projection=out_proj, rnn_cell=cell,
                            conditional_gru=conditional, encoder_projection=enc_proj, rnn_size=7, mlp_layers=2)
    enc = {"temporal_states": _t(pre + "states"), "temporal_mask": _t(pre + "mask"), "output": _t(pre + "enc_out")}
    gold = torch.from_numpy(G[pre + "gold"])
    init = O.decoder_initial_state(p, spec, enc["output"], enc, bsz=3)
    assert np.abs(init.numpy() - G[pre + "initial_state"]).max() < 2e-6
    train = O.decoder_train(p, spec, enc, gold)
    assert np.abs(train["train_logits"].numpy() - G[pre + "train_logits"]).max() < 1e-5
    assert np.abs(train["rnn_outputs"].numpy() - G[pre + "train_rnn_outputs"]).max() < 1e-5
    assert abs(float(train["train_loss"]) - float(G[pre + "train_loss"])) < 1e-5
    run = O.decoder_greedy(p, spec, enc)
    assert np.array_equal(run["output_symbols"].numpy(), G[pre + "run_symbols"])
    assert np.abs(run["runtime_logits"].numpy() - G[pre + "run_logits"]).max() < 1e-5
    # every dense layer the reference asked for is a parameter the oracle read, under that name
    assert {n for n in G[pre + "dense_names"].tolist()} <= set(p)


def test_attention_on_input_cannot_be_built_in_the_reference():
    """Decoder.input_plus_attention reads `feedables.prev_contexts` (decoders/decoder.py:273); the
    contexts live in `feedables.other`, so attention_on_input=True raises while the graph is built.
    The product therefore rejects the option instead of guessing a behaviour."""
    assert "prev_contexts" in str(G["attention_on_input_error"])


@pytest.mark.parametrize("strategy", ["serial", "parallel", "flat", "hierarchical"])
def test_multi_source_cross_attention_strategies(strategy):
    """TransformerDecoder.layer over TWO encoders with each attention_combination_strategy
    (attention/transformer_cross_layer.py:12-263): scopes enc_<i> / enc_hier, where the LayerNorms sit,
    what the residual adds, the time-axis concatenation of `flat`, the [B*T, n, d] second attention of
    `hierarchical`.  The variables were drawn when the reference asked for them, so their names are the
    reference's."""
    name = "tms_" + strategy
    p = {k[4:]: _t(k) for k in G.files if k.startswith("mv::" + name + "/")}
    spec = O.TransformerDecoderSpec(name, 2, 3, 3, 9)
    states = O.transformer_decoder_stack(
        p, spec, _t("ms_in"), _t("ms_mask"), [_t("ms_enc_a"), _t("ms_enc_b")], [_t("ms_mask_a"), _t("ms_mask_b")],
        strategy=strategy, heads_enc=[3, 3] if strategy == "flat" else [3, 2],
        heads_hier=4 if strategy == "hierarchical" else None)
    assert np.abs(states.numpy() - G[name + "_states"]).max() < 3e-5
    scopes = {k.split("/encdec_attention/")[1].split("/")[0] for k in p
              if "/layer_0/encdec_attention/" in k and k.endswith("/kernel")}
    assert scopes == {"serial": {"enc_0", "enc_1"}, "parallel": {"enc_0", "enc_1"},
                      "flat": {"keys_proj", "output_proj", "query_proj", "vals_proj"},
                      "hierarchical": {"enc_0", "enc_1", "enc_hier"}}[strategy]


def test_transformer_training_logits_ignore_supress_unk():
    """supress_unk lives in get_body's state_to_logits (autoregressive.py:454-457): the Transformer's
    training pass computes its logits itself and carries no -1e9 <unk> column, its run-time loop does."""
    p, spec, enc = _transformer_setup()
    spec.supress_unk = True
    gold = torch.from_numpy(G["tloop_gold"]).t().contiguous()
    train = O.transformer_decoder_train(p, spec, enc, gold)
    assert np.abs(train["logits"].transpose(0, 1).numpy() - G["tloop_unk_train_logits"]).max() < 5e-5
    assert np.array_equal(G["tloop_unk_train_logits"], G["tloop_train_logits"])
    run = O.transformer_decoder_greedy(p, spec, enc)
    keep = np.arange(G["tloop_unk_run_logits"].shape[-1]) != O.UNK
    assert np.abs(run["logits"].numpy()[..., keep] - G["tloop_unk_run_logits"][..., keep]).max() < 5e-5
    assert float(G["tloop_unk_run_logits"][..., O.UNK].max()) < -1e8 and float(run["logits"][..., O.UNK].max()) < -1e8


@pytest.mark.parametrize("tag,heads", [("h3", 3), ("h1", 1)])
def test_rnn_decoder_with_scaled_dot_attention_objects(tag, heads):
    """The reference's Decoder run whole with a MultiHeadAttention (keys and values from different tensors) and a
    ScaledDotProdAttention as its attentions (attention/scaled_dot_product.py:246-402, tests/post-edit.ini):
    the oracle's decoder over a LIST of attention objects reproduces its training pass, greedy loop and per-head
    histories; the head projections the run created live in the DECODER's step scope."""
    dname, pre = "md_" + tag, "md_{}_".format(tag)
    p = {k[4:]: _t(k) for k in G.files if k.startswith("mv::" + dname + "/")}
    p[dname + "/word_embeddings"] = _t(pre + "table")
    p[dname + "/state_to_word_W"], p[dname + "/state_to_word_b"] = _t(pre + "w"), _t(pre + "b")
    want_dense = ["{}/attention_decoder/dense/kernel".format(dname), "{}/initial_state/encoders_projection/kernel".format(dname)]
    if heads > 1:
        want_dense += ["{}/attention_decoder/{}_proj/kernel".format(dname, n) for n in ("keys", "output", "query", "vals")]
    assert sorted(G[pre + "dense_names"].tolist()) == sorted(want_dense)
    assert G[pre + "context_sizes"].tolist() == [12, 12]
    keys, values, mask = _t(pre + "keys"), _t(pre + "values"), _t(pre + "mask")
    scope = dname + "/attention_decoder"
    attend = [lambda q: O.multihead_attention_step(p, scope, q, keys, values, mask, heads),
              lambda q: O.multihead_attention_step(p, scope, q, keys, keys, mask, 1)]
    spec = O.RNNDecoderSpec(dname, None, max_output_len=5, output_projection="tanh")
    enc = {"output": torch.cat([_t(pre + "enc_out0"), _t(pre + "enc_out1")], 1)}
    gold = torch.from_numpy(G[pre + "gold"])
    train = O.decoder_train(p, spec, enc, gold, attend=attend)
    assert np.abs(train["train_logits"].numpy() - G[pre + "train_logits"]).max() < 1e-5
    assert np.abs(train["rnn_outputs"].numpy() - G[pre + "train_rnn_outputs"]).max() < 1e-5
    assert abs(float(train["train_loss"]) - float(G[pre + "train_loss"])) < 1e-5
    for i in range(heads):
        assert np.abs(train["attention_weights"][0][:, :, i].numpy() - G["{}train_mha_head{}".format(pre, i)]).max() < 1e-6
    assert np.abs(train["attention_weights"][1][:, :, 0].numpy() - G[pre + "train_sdp_head0"]).max() < 1e-6
    run = O.decoder_greedy(p, spec, enc, attend=attend)
    assert np.array_equal(run["output_symbols"].numpy(), G[pre + "run_symbols"])
    assert np.abs(run["runtime_logits"].numpy() - G[pre + "run_logits"]).max() < 1e-5
    keys_want = sorted("{}_{}_head{}".format(dname, m, i) for m in ("train", "run") for i in range(heads))
    assert G[pre + "history_keys"].tolist() == keys_want + sorted("{}_{}_head0".format(dname, m) for m in ("train", "run"))
