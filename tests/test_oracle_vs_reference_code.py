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
