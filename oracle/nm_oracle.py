"""CPU oracle: an op-for-op restatement of the Neural Monkey hot path (torch-CPU).

TEST INFRASTRUCTURE ONLY.  Nothing under neuralmonkey_b200/ may import this
module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs do, and only as the checker or the timed CPU baseline.

PARITY: pinned to the reference's own code; the TensorFlow-library primitives are restated.
The reference (ufal/neuralmonkey @ 8b14652) evaluates this path with TensorFlow
(`tensorflow>=1.12.0,<1.13`, requirements.txt:13), a third-party dependency that is
absent from /root/reference and cannot be installed here (Python 3.12, no network);
its own tests hold no numeric golden vectors for the path.  So:

  * everything the REFERENCE implements - the attention functions and objects, whole
    encoder / decoder stacks, the decoding loops, beam search, the loss tensors, the
    trainer's host logic, the host pipeline - is pinned by running the reference's own
    Python, imported from /root/reference, over a numpy stand-in for the TF ops it
    calls (tests/golden/tf_numpy_shim.py); the generated vectors and the generating
    scripts are committed (tests/golden/make_*_golden.py) and this file is tested
    against them (tests/test_oracle_vs_reference_code.py, tests/test_host_golden.py);
  * what TensorFlow itself implements is restated below from its published algorithm,
    with parity anchored on the reference's call sites (scopes, variable names, what is
    fed to which cell), which the stand-in runs pin:

  * tf.contrib.rnn.GRUCell (TF 1.12 rnn_cell_impl.py):
        gate_inputs = matmul(concat([x, h], 1), gates/kernel) + gates/bias
        r, u = split(sigmoid(gate_inputs), 2)
        candidate = matmul(concat([x, r * h], 1), candidate/kernel) + candidate/bias
        c = tanh(candidate);  new_h = u * h + (1 - u) * c
  * tf.nn.dynamic_rnn with sequence_length: for t >= length the output is zero and
    the state is copied through; bidirectional_dynamic_rnn runs the backward cell on
    tf.reverse_sequence(x, lengths) and reverses its outputs back.
  * tf.nn.softmax / log_softmax over the last axis; tf.argmax and tf.nn.top_k
    return the lowest index among equal values.
  * tf.contrib.seq2seq.sequence_loss(average_across_* = False):
        sparse_softmax_cross_entropy_with_logits(targets, logits) * weights
  * tf.train.AdamOptimizer:  lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)
        m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; var -= lr_t * m / (sqrt(v) + eps)
  * tf.clip_by_norm(t, c) = t * c / max(||t||_2, c)

Everything runs in the dtype of the parameters passed in (float32 to mirror the
reference, float64 to measure kernel error); gradients come from torch.autograd
over the same restated graph.
"""
import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

PAD, START, END, UNK = 0, 1, 2, 3   # neuralmonkey/vocabulary.py:19-29
INF = 1e9                           # decoders/beam_search_decoder.py:43

Params = Dict[str, torch.Tensor]


# ---------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------
def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
               eps: float = 1e-6) -> torch.Tensor:
    """neuralmonkey/tf_utils.py:189-219."""
    mean = x.mean(dim=-1, keepdim=True)
    variance = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    norm_x = (x - mean) * torch.rsqrt(variance + eps)
    return norm_x * gamma + beta


def gru_cell(x: torch.Tensor, h: torch.Tensor, wg: torch.Tensor, bg: torch.Tensor,
             wc: torch.Tensor, bc: torch.Tensor) -> torch.Tensor:
    """tf.contrib.rnn.GRUCell as used by OrthoGRUCell (nn/ortho_gru_cell.py:44-53)."""
    gates = torch.sigmoid(torch.cat([x, h], 1) @ wg + bg)
    size = h.shape[1]
    r, u = gates[:, :size], gates[:, size:]
    c = torch.tanh(torch.cat([x, r * h], 1) @ wc + bc)
    return u * h + (1 - u) * c


def _dense(p: Params, scope: str, x: torch.Tensor) -> torch.Tensor:
    """tf.layers.dense under `scope`; the bias is used when the layer has one."""
    y = x @ p[scope + "/kernel"]
    return y + p[scope + "/bias"] if scope + "/bias" in p else y


def nematus_gru_cell(p: Params, scope: str, x: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """NematusGRUCell.call (nn/ortho_gru_cell.py:72-105): separate input and state projections, the
    reset gate applied AFTER the state projection of the candidate:
        [r, u] = sigmoid(state_proj_g(h) + input_proj_g(x));  c = tanh(state_proj_c(h) * r + input_proj_c(x))
        h' = u * h + (1 - u) * c
    Which of the projections carry a bias (use_state_bias / use_input_bias) shows in the parameters."""
    gates = torch.sigmoid(_dense(p, scope + "gates/state_proj", h) + _dense(p, scope + "gates/input_proj", x))
    size = h.shape[1]
    r, u = gates[:, :size], gates[:, size:]
    c = torch.tanh(_dense(p, scope + "candidate/state_proj", h) * r + _dense(p, scope + "candidate/input_proj", x))
    return u * h + (1 - u) * c


def lstm_cell(p: Params, scope: str, x: torch.Tensor, c: torch.Tensor, h: torch.Tensor):
    """tf.nn.rnn_cell.LSTMCell with its defaults (no peepholes, forget_bias = 1): gate order i, j, f, o
    in one [in + H, 4H] kernel; returns (c', h')."""
    z = torch.cat([x, h], 1) @ p[scope + "kernel"] + p[scope + "bias"]
    i, j, f, o = z.chunk(4, dim=1)
    new_c = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
    return new_c, torch.sigmoid(o) * torch.tanh(new_c)


def dynamic_rnn(cell: Callable, size: int, x: torch.Tensor, lengths: Optional[torch.Tensor],
                h0: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """tf.nn.dynamic_rnn(cell, x, sequence_length=lengths) for a cell `h' = cell(x_t, h)` whose output
    is its state: past a sentence's length the output is zero and the state is carried."""
    bsz, steps, _ = x.shape
    h = h0 if h0 is not None else x.new_zeros(bsz, size)
    outs = []
    for t in range(steps):
        new_h = cell(x[:, t], h)
        if lengths is not None:
            live = (t < lengths).to(x.dtype).unsqueeze(1)
            outs.append(new_h * live)
            h = new_h * live + h * (1 - live)
        else:
            outs.append(new_h)
            h = new_h
    return torch.stack(outs, 1), h


def dynamic_gru(x: torch.Tensor, lengths: Optional[torch.Tensor], wg, bg, wc, bc,
                h0: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """tf.nn.dynamic_rnn(GRUCell, x, sequence_length=lengths)."""
    return dynamic_rnn(lambda xt, h: gru_cell(xt, h, wg, bg, wc, bc), wc.shape[1], x, lengths, h0)


def reverse_sequence(x: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """tf.reverse_sequence(x, lengths, seq_axis=1)."""
    out = x.clone()
    for b in range(x.shape[0]):
        n = int(lengths[b])
        if n > 0:
            out[b, :n] = x[b, :n].flip(0)
    return out


def bidirectional_gru(x, lengths, fw: Tuple, bw: Tuple) -> Tuple[torch.Tensor, torch.Tensor]:
    """rnn_layer(..., direction='bidirectional') (encoders/recurrent.py:82-95)."""
    out_fw, fin_fw = dynamic_gru(x, lengths, *fw)
    out_bw_rev, fin_bw = dynamic_gru(reverse_sequence(x, lengths), lengths, *bw)
    out_bw = reverse_sequence(out_bw_rev, lengths)
    return torch.cat([out_fw, out_bw], 2), torch.cat([fin_fw, fin_bw], 1)


def sentence_mask(ids: torch.Tensor, dtype) -> torch.Tensor:
    """vocabulary.sentence_mask (vocabulary.py:357-358): id != PAD as float."""
    return (ids != PAD).to(dtype)


# ---------------------------------------------------------------------------
# SentenceEncoder (encoders/recurrent.py:113-314, model/sequence.py:170-199)
# ---------------------------------------------------------------------------
def embedded_sequence(p: Params, name: str, factor_ids: Sequence[torch.Tensor],
                      scale_embeddings_by_depth: bool = False) -> Dict[str, torch.Tensor]:
    """EmbeddedFactorSequence.temporal_states / temporal_mask (model/sequence.py:170-199): every factor
    is looked up in its own `embedding_matrix_<i>`, optionally scaled by sqrt(size), multiplied by the
    mask of the FIRST factor, and the factors are concatenated on the feature axis."""
    dtype = p[name + "/embedding_matrix_0"].dtype
    mask = sentence_mask(factor_ids[0], dtype)
    factors = []
    for i, ids in enumerate(factor_ids):
        matrix = p["{}/embedding_matrix_{}".format(name, i)]
        emb = matrix[ids]
        if scale_embeddings_by_depth:
            emb = emb * matrix.shape[-1] ** 0.5
        factors.append(emb * mask.unsqueeze(-1))
    return {"temporal_states": torch.cat(factors, 2), "temporal_mask": mask}


def _cell_params(p: Params, scope: str) -> Tuple:
    return tuple(p[scope + n] for n in ("gates/kernel", "gates/bias", "candidate/kernel", "candidate/bias"))


def rnn_layer(p: Params, scope: str, x: torch.Tensor, lengths: torch.Tensor, direction: str,
              cell_type: str = "GRU", size: int = 0):
    """rnn_layer (encoders/recurrent.py:71-110); variable scopes as TensorFlow names them:
    <scope>/bidirectional_rnn/{fw,bw}/<cell> or <scope>/rnn/<cell>, <cell> = OrthoGRUCell (the scope
    the reference passes, nn/ortho_gru_cell.py:51) or nematus_gru_cell (TensorFlow's default name for
    the layer class; `size` is needed for that cell only)."""
    def run(cell_scope, inputs):
        if cell_type == "LSTM":      # state (c, h), both carried past the length; the layer's final state is h
            cell_scope += "lstm_cell/"
            c = h = inputs.new_zeros(inputs.shape[0], size)
            outs = []
            for t in range(inputs.shape[1]):
                new_c, new_h = lstm_cell(p, cell_scope, inputs[:, t], c, h)
                live = (t < lengths).to(inputs.dtype).unsqueeze(1)
                outs.append(new_h * live)
                c, h = new_c * live + c * (1 - live), new_h * live + h * (1 - live)
            return torch.stack(outs, 1), h
        if cell_type == "NematusGRU":
            cell_scope += "nematus_gru_cell/"
            return dynamic_rnn(lambda xt, h: nematus_gru_cell(p, cell_scope, xt, h), size, inputs, lengths)
        return dynamic_gru(inputs, lengths, *_cell_params(p, cell_scope + "OrthoGRUCell/"))

    if direction == "bidirectional":
        out_fw, fin_fw = run(scope + "/bidirectional_rnn/fw/", x)
        out_bw_rev, fin_bw = run(scope + "/bidirectional_rnn/bw/", reverse_sequence(x, lengths))
        return (torch.cat([out_fw, reverse_sequence(out_bw_rev, lengths)], 2), torch.cat([fin_fw, fin_bw], 1))
    if direction == "backward":
        out_rev, final = run(scope + "/rnn/", reverse_sequence(x, lengths))
        return reverse_sequence(out_rev, lengths), final
    return run(scope + "/rnn/", x)


def recurrent_encoder(p: Params, prefix: str, inputs: torch.Tensor, mask: torch.Tensor,
                      rnn_layers: Sequence[Tuple] = ((0, "bidirectional"),),
                      add_residual: bool = False, add_layer_norm: bool = False,
                      include_final_layer_norm: bool = True) -> Dict[str, torch.Tensor]:
    """RecurrentEncoder.rnn (encoders/recurrent.py:180-218), dropout off.  Layer i lives in scope
    rnn_<i>_<direction>; with add_layer_norm its INPUT is normalised first (own LayerNorm variables)
    and - note - the residual then adds the normalised input, not the raw one; the residual applies
    only when input and output widths agree; the final LayerNorm normalises the states and the final
    state with the SAME variables (:215-216)."""
    lengths = mask.sum(1).to(torch.int64)             # model/stateful.py:56-62
    layer_input, layer_final = inputs, inputs[:, -1]
    for i, layer in enumerate(rnn_layers):
        size, direction = layer[0], layer[1]
        cell_type = layer[2] if len(layer) > 2 else "GRU"
        scope = "{}/rnn_{}_{}".format(prefix, i, direction)
        if add_layer_norm:
            layer_input = layer_norm(layer_input, p[scope + "/LayerNorm/gamma"], p[scope + "/LayerNorm/beta"])
        layer_output, layer_final_output = rnn_layer(p, scope, layer_input, lengths, direction, cell_type, size)
        if add_residual and layer_input.shape[-1] == layer_output.shape[-1]:
            layer_input = layer_input + layer_output
            layer_final = layer_final + layer_final_output
        else:
            layer_input, layer_final = layer_output, layer_final_output
    if include_final_layer_norm:
        gamma, beta = p[prefix + "/LayerNorm/gamma"], p[prefix + "/LayerNorm/beta"]
        layer_input, layer_final = layer_norm(layer_input, gamma, beta), layer_norm(layer_final, gamma, beta)
    return {"temporal_states": layer_input, "output": layer_final, "temporal_mask": mask}


def sentence_encoder(p: Params, prefix: str, ids: torch.Tensor) -> Dict[str, torch.Tensor]:
    """SentenceEncoder (recurrent.py:221-314): one embedded factor, one bidirectional GRU layer."""
    seq = embedded_sequence(p, prefix + "_input", [ids])
    return recurrent_encoder(p, prefix, seq["temporal_states"], seq["temporal_mask"])


# ---------------------------------------------------------------------------
# Attention (attention/feed_forward.py:47-166)
# ---------------------------------------------------------------------------
def bahdanau_precompute(p: Params, prefix: str, states: torch.Tensor) -> torch.Tensor:
    """hidden_features: 1x1 conv == states @ key_projection (feed_forward.py:111-118)."""
    return states @ p[prefix + "/attn_key_projection"]


def bahdanau_step(p: Params, prefix: str, query: torch.Tensor, hidden: torch.Tensor,
                  states: torch.Tensor, mask: Optional[torch.Tensor]):
    """Attention.attention (feed_forward.py:125-166)."""
    y = query @ p[prefix + "/Attention/attn_query_projection"] + p[prefix + "/attn_projection_bias"]
    v = p[prefix + "/attn_similarity_v"]
    energies = (v * torch.tanh(hidden + y.unsqueeze(1))).sum(-1) + p[prefix + "/attn_bias"]
    if mask is None:
        weights = torch.softmax(energies, dim=-1)
    else:
        weights_all = torch.softmax(energies, dim=-1) * mask
        norm = weights_all.sum(1, keepdim=True) + 1e-8
        weights = weights_all / norm
    context = (weights.unsqueeze(-1) * states).sum(1)
    return context, weights


# ---------------------------------------------------------------------------
# Decoder (decoders/decoder.py:226-358, autoregressive.py:381-519)
# ---------------------------------------------------------------------------
class RNNDecoderSpec:
    """Static configuration of decoders.decoder.Decoder for the oracle."""

    def __init__(self, prefix: str, att_prefix: str, max_output_len: int,
                 output_projection: str = "tanh", supress_unk: bool = False, rnn_cell: str = "GRU",
                 conditional_gru: bool = False, encoder_projection: str = "linear",
                 rnn_size: Optional[int] = None, mlp_layers: int = 0) -> None:
        self.prefix = prefix
        self.att_prefix = att_prefix
        self.max_output_len = max_output_len
        self.output_projection = output_projection  # "tanh" | "maxout" | "nematus" | "mlp"
        self.supress_unk = supress_unk
        self.rnn_cell = rnn_cell                    # "GRU" | "NematusGRU"
        self.conditional_gru = conditional_gru
        self.encoder_projection = encoder_projection  # "linear" | "nematus" | "concat" | "empty"
        self.rnn_size = rnn_size                    # only read by the "empty" projection
        self.mlp_layers = mlp_layers


def decoder_initial_state(p: Params, spec: RNNDecoderSpec, enc_output: Optional[torch.Tensor],
                          enc: Optional[Dict[str, torch.Tensor]] = None, bsz: int = 0) -> torch.Tensor:
    """Decoder.initial_state (decoder.py:226-251) over the encoder projections
    (encoder_projection.py:30-145); dropout off.
        linear   dense(concatenated encoder outputs)                      (:47-73)
        concat   the concatenated encoder outputs themselves             (:76-96)
        nematus  tanh(dense(mask-weighted MEAN of the temporal states)) (:99-145)
        empty    zeros(rnn_size), tiled over the batch                   (:30-44, decoder.py:245-250)"""
    scope = spec.prefix + "/initial_state/encoders_projection"
    if spec.encoder_projection == "linear":
        return _dense(p, scope, enc_output)
    if spec.encoder_projection == "concat":
        return enc_output
    if spec.encoder_projection == "nematus":
        mask = enc["temporal_mask"]
        means = (enc["temporal_states"] * mask.unsqueeze(2)).sum(1) / mask.sum(1, keepdim=True)
        return torch.tanh(_dense(p, scope, means))
    if spec.encoder_projection == "empty":
        return torch.zeros(bsz, spec.rnn_size)
    raise ValueError(spec.encoder_projection)


def output_projection(p: Params, spec: RNNDecoderSpec, cell_output, embedded_input, context):
    """nonlinear_output / maxout_output / nematus_output / mlp_output
    (output_projection.py:76-188, nn/projection.py:7-57)."""
    step = spec.prefix + "/attention_decoder/"
    if spec.output_projection == "nematus":
        # three separate projections summed, then tanh (:76-112)
        return torch.tanh(_dense(p, step + "rnn_state", cell_output) + _dense(p, step + "prev_out", embedded_input)
                          + _dense(p, step + "context", context))
    cat = torch.cat([cell_output, embedded_input, context], 1)
    if spec.output_projection == "mlp":
        # multilayer_projection: the activation follows EVERY layer, the last one too (:163-188)
        for i in range(spec.mlp_layers):
            cat = torch.tanh(_dense(p, "{}deep_output_mlp/mlp_layer_{}".format(step, i), cat))
        return cat
    if spec.output_projection == "tanh":
        return torch.tanh(_dense(p, step + "dense", cat))
    z = _dense(p, step + "MaxoutProjection/MaxoutProjection", cat)
    size = z.shape[1] // 2
    # reshape [-1,1,2,size] + max_pool over the length-2 axis: first half vs second half
    return torch.maximum(z[:, :size], z[:, size:])


def state_to_logits(p: Params, spec: RNNDecoderSpec, state: torch.Tensor) -> torch.Tensor:
    """autoregressive.py:450-459."""
    logits = state @ p[spec.prefix + "/state_to_word_W"] + p[spec.prefix + "/state_to_word_b"]
    if spec.supress_unk:
        unk_mask = torch.zeros(logits.shape[1], dtype=logits.dtype)
        unk_mask[UNK] = -1e9
        logits = logits + unk_mask
    return logits


def _decoder_cell(p: Params, spec: RNNDecoderSpec, scope: str, x: torch.Tensor, h: torch.Tensor):
    if spec.rnn_cell == "NematusGRU":
        return nematus_gru_cell(p, scope, x, h)
    return gru_cell(x, h, p[scope + "gates/kernel"], p[scope + "gates/bias"],
                    p[scope + "candidate/kernel"], p[scope + "candidate/bias"])


def multihead_attention_step(p: Params, scope: str, query: torch.Tensor, keys: torch.Tensor,
                             values: torch.Tensor, keys_mask: Optional[torch.Tensor], heads: int):
    """MultiHeadAttention.attention / ScaledDotProdAttention (scaled_dot_product.py:296-350): the attention
    OBJECT an RNN decoder queries once per step - `attention()` over a one-step query sequence; the head
    projections (heads > 1) are created while the decoder's step scope is open, i.e. `scope` is
    `<decoder>/attention_decoder`.  Returns (context [B, dim], weights [B, heads, time])."""
    context, weights = multihead_attention(p, scope, query.unsqueeze(1), keys, values, keys_mask, heads)
    return context[:, 0], weights[:, :, 0]


def decoder_step(p: Params, spec: RNNDecoderSpec, embedded_input, prev_output, hidden, states, mask,
                 attend: Optional[Sequence[Callable]] = None):
    """Decoder.next_state, GRU / NematusGRU branch, attention_on_input=False (decoder.py:279-358; the
    reference cannot build attention_on_input=True - `feedables.prev_contexts`, :273, does not exist);
    dropout off so prev_rnn_output == cell_output.  With `conditional_gru` the context is run through
    a second cell (scope `cond_gru_2_cell`) whose state is the first cell's output, and ITS output is
    what the projection, the history and the next step see - the attention was queried with the first
    cell's output (:303-325)."""
    step = spec.prefix + "/attention_decoder/"
    if spec.rnn_cell == "LSTM":
        # the LSTM branch (:326-339): the loop carries (prev_rnn_state = c, prev_rnn_output = h), both
        # initialised with the initial state; `prev_output` is that pair here, and so is the second
        # return value.  No conditional cell in this branch.
        prev_c, prev_h = prev_output if isinstance(prev_output, tuple) else (prev_output, prev_output)
        new_c, cell_output = lstm_cell(p, step + "lstm_cell/", embedded_input, prev_c, prev_h)
        if attend is not None:
            attended = [f(cell_output) for f in attend]
            context, weights = torch.cat([c for c, _w in attended], 1), [w for _c, w in attended]
        else:
            context, weights = bahdanau_step(p, spec.att_prefix, cell_output, hidden, states, mask)
        output = output_projection(p, spec, cell_output, embedded_input, context)
        return output, (new_c, cell_output), context, weights
    first = step + ("nematus_gru_cell/" if spec.rnn_cell == "NematusGRU" else "OrthoGRUCell/")
    cell_output = _decoder_cell(p, spec, first, embedded_input, prev_output)
    if attend is not None:
        # any list of attention objects (decoder.py:290-300): each is queried with the cell output, the
        # contexts are concatenated in the order of `attentions` wherever they are consumed
        attended = [f(cell_output) for f in attend]
        context = torch.cat([c for c, _w in attended], 1)
        weights = [w for _c, w in attended]
    else:
        context, weights = bahdanau_step(p, spec.att_prefix, cell_output, hidden, states, mask)
    if spec.conditional_gru:
        cell_output = _decoder_cell(p, spec, step + "cond_gru_2_cell/", context, cell_output)
    output = output_projection(p, spec, cell_output, embedded_input, context)
    return output, cell_output, context, weights


def autoregressive_loop(next_output: Callable, to_logits: Callable, embed: Callable, bsz: int,
                        max_len: int, train_inputs: Optional[torch.Tensor] = None) -> Dict[str, list]:
    """AutoregressiveDecoder.decoding_loop + get_body (autoregressive.py:425-562), the part every
    decoder shares.  `next_output(embedded_input, finished)` is the subclass's next_state (it keeps its
    own recurrent state), `to_logits` the vocabulary projection, `embed` the embedding lookup.

    Per step: logits = to_logits(output); next symbol = gold `train_inputs[step]` (training) or
    argmax over the FULL vocabulary (runtime); `symbol *= not finished` (pad once finished);
    `finished |= symbol == </s>`; the mask appended is `not finished` AFTER the update.  The loop
    runs while `not all(finished) and step < max_len` (:425-437)."""
    finished = torch.zeros(bsz, dtype=torch.bool)
    embedded = embed(torch.full((bsz,), START, dtype=torch.int64))
    hist = {"logits": [], "outputs": [], "symbols": [], "mask": [], "extra": []}
    step = 0
    while (not bool(finished.all())) and step < max_len:
        output, extra = next_output(embedded, finished)
        logits = to_logits(output)
        chosen = train_inputs[step] if train_inputs is not None else torch.argmax(logits, dim=1)
        symbols = chosen * (~finished).to(torch.int64)
        finished = finished | (symbols == END)
        embedded = embed(symbols)
        hist["logits"].append(logits)
        hist["outputs"].append(output)
        hist["symbols"].append(symbols)
        hist["mask"].append(~finished)
        hist["extra"].append(extra)
        step += 1
    return hist


def sequence_xents(logits: torch.Tensor, targets: torch.Tensor, mask: torch.Tensor,
                   label_smoothing: Optional[float] = None) -> torch.Tensor:
    """train_xents (autoregressive.py:292-310) for logits [.., V], targets and mask of the leading shape.

    Without smoothing: per-position cross-entropy times the mask.  WITH `label_smoothing` the reference
    hands sequence_loss a function built on `tf.losses.softmax_cross_entropy`, which smooths the labels to
    onehot * (1 - s) + s / V but then REDUCES to one scalar - the mean over ALL positions, padding
    included, their target being <pad> - and sequence_loss multiplies that scalar by the mask.  So every
    unmasked position carries the same number, and train_loss (sum / sum(mask)) is that number."""
    logprobs = torch.log_softmax(logits, dim=-1)
    if not label_smoothing:
        return -logprobs.gather(-1, targets.unsqueeze(-1)).squeeze(-1) * mask
    vocab = logits.shape[-1]
    smoothed = torch.nn.functional.one_hot(targets, vocab).to(logits.dtype) * (1.0 - label_smoothing) \
        + label_smoothing / vocab
    return -(smoothed * logprobs).sum(-1).mean() * mask


def decoder_train(p: Params, spec: RNNDecoderSpec, enc: Dict[str, torch.Tensor],
                  train_inputs: torch.Tensor, label_smoothing: Optional[float] = None,
                  attend: Optional[Sequence[Callable]] = None) -> Dict[str, torch.Tensor]:
    """decoding_loop(train_mode=True) + train_xents/train_loss (autoregressive.py:292-316,532-562).

    train_inputs: [T, B] int64 = padded references with </s> appended (feed_dict :579-582),
    already transposed to time-major as `train_inputs` is (:199-202)."""
    emb = p[spec.prefix + "/word_embeddings"]
    _steps, bsz = train_inputs.shape
    states, mask = enc.get("temporal_states"), enc.get("temporal_mask")
    hidden = bahdanau_precompute(p, spec.att_prefix, states) if attend is None else None
    rnn = {"prev": decoder_initial_state(p, spec, enc.get("output"), enc, bsz)}

    def next_output(embedded, _finished):
        output, rnn["prev"], _ctx, w = decoder_step(p, spec, embedded, rnn["prev"], hidden, states, mask, attend)
        return output, (w, rnn["prev"][1] if isinstance(rnn["prev"], tuple) else rnn["prev"])

    hist = autoregressive_loop(next_output, lambda o: state_to_logits(p, spec, o), lambda ids: emb[ids], bsz,
                               spec.max_output_len, train_inputs)
    logits_hist, out_states = hist["logits"], hist["outputs"]
    att_weights, rnn_outputs = [e[0] for e in hist["extra"]], [e[1] for e in hist["extra"]]
    logits_t = torch.stack(logits_hist, 0)                       # [T,B,V]
    train_mask = sentence_mask(train_inputs, logits_t.dtype)     # [T,B]
    xents = sequence_xents(logits_t, train_inputs[:logits_t.shape[0]], train_mask[:logits_t.shape[0]],
                           label_smoothing)
    loss = xents.sum() / train_mask.sum()
    return {"train_logits": logits_t, "train_xents": xents.t(), "train_loss": loss,
            "train_mask": train_mask, "train_output_states": torch.stack(out_states, 0),
            "attention_weights": (torch.stack(att_weights, 0) if attend is None else
                                  [torch.stack([w[i] for w in att_weights], 0) for i in range(len(attend))]),
            "rnn_outputs": torch.stack(rnn_outputs, 0)}


def decoder_greedy(p: Params, spec: RNNDecoderSpec, enc: Dict[str, torch.Tensor],
                   train_inputs: Optional[torch.Tensor] = None,
                   attend: Optional[Sequence[Callable]] = None) -> Dict[str, torch.Tensor]:
    """decoding_loop(train_mode=False): argmax feedback (autoregressive.py:446-519) and, when
    references are given, runtime_xents / runtime_loss (:351-371)."""
    emb = p[spec.prefix + "/word_embeddings"]
    states, mask = enc.get("temporal_states"), enc.get("temporal_mask")
    bsz = states.shape[0] if states is not None else enc["output"].shape[0]
    hidden = bahdanau_precompute(p, spec.att_prefix, states) if attend is None else None
    rnn = {"prev": decoder_initial_state(p, spec, enc.get("output"), enc, bsz)}

    def next_output(embedded, _finished):
        output, rnn["prev"], _ctx, _w = decoder_step(p, spec, embedded, rnn["prev"], hidden, states, mask, attend)
        return output, None

    hist = autoregressive_loop(next_output, lambda o: state_to_logits(p, spec, o), lambda ids: emb[ids], bsz,
                               spec.max_output_len)
    logits_hist, symbols, out_mask = hist["logits"], hist["symbols"], hist["mask"]
    logits_t = torch.stack(logits_hist, 0)
    res = {"runtime_logits": logits_t, "runtime_logprobs": torch.log_softmax(logits_t, -1),
           "output_symbols": torch.stack(symbols, 0), "runtime_mask": torch.stack(out_mask, 0),
           "decoded": torch.argmax(logits_t[:, :, 1:], -1) + 1}
    if train_inputs is not None:
        min_time = min(train_inputs.shape[0], logits_t.shape[0])
        lp = torch.log_softmax(logits_t[:min_time], -1)
        tmask = sentence_mask(train_inputs, logits_t.dtype)[:min_time]
        xents = -lp.gather(2, train_inputs[:min_time].unsqueeze(-1)).squeeze(-1) * tmask
        res["runtime_xents"] = xents.t()
        res["runtime_loss"] = xents.sum() / res["runtime_mask"].to(logits_t.dtype).sum()
    return res


# ---------------------------------------------------------------------------
# Transformer (encoders/transformer.py, decoders/transformer.py,
# attention/scaled_dot_product.py, attention/transformer_cross_layer.py)
# ---------------------------------------------------------------------------
def position_signal(dimension: int, length: int, dtype=torch.float32) -> torch.Tensor:
    """encoders/transformer.py:23-45 - [1, length, dimension]; sin half then cos half."""
    positions = torch.arange(length, dtype=torch.float32)
    num_timescales = dimension // 2
    log_timescale_increment = math.log(1.0e4) / (num_timescales - 1)
    inv_timescales = torch.exp(torch.arange(num_timescales, dtype=torch.float32)
                               * -log_timescale_increment)
    scaled_time = positions.unsqueeze(1) * inv_timescales.unsqueeze(0)
    signal = torch.cat([torch.sin(scaled_time), torch.cos(scaled_time)], dim=1)
    if dimension % 2:
        signal = torch.nn.functional.pad(signal, (0, 1))
    return signal.reshape(1, length, dimension).to(dtype)


def _split_for_heads(x: torch.Tensor, heads: int, head_dim: int) -> torch.Tensor:
    """scaled_dot_product.py:24-41: [B,T,D] -> [B,heads,T,head_dim]."""
    return x.reshape(x.shape[0], x.shape[1], heads, head_dim).permute(0, 2, 1, 3)


def multihead_attention(p: Params, scope: str, queries, keys, values, keys_mask, heads: int,
                        masked: bool = False, use_bias: bool = False, drop_mask=None):
    """attention() (scaled_dot_product.py:98-226).  The dropout callback is the identity, or - with
    `drop_mask` [B, heads, Tq, Tk] holding 0 or 1/keep_prob - a multiplication of the softmax weights
    by that mask before they are applied to the values (:208-214)."""
    dim = queries.shape[-1]
    head_dim = dim // heads

    def dense(x, name):
        y = x @ p["{}/{}/kernel".format(scope, name)]
        return y + p["{}/{}/bias".format(scope, name)] if use_bias else y

    if heads > 1:
        queries, keys, values = dense(queries, "query_proj"), dense(keys, "keys_proj"), dense(values, "vals_proj")
    q = _split_for_heads(queries / math.sqrt(head_dim), heads, head_dim)
    k = _split_for_heads(keys, heads, head_dim)
    v = _split_for_heads(values, heads, head_dim)
    energies = q @ k.transpose(-1, -2)
    if masked:  # mask_future (:44-66): lower triangle kept, the rest REPLACED by -1e9
        tq, tk = energies.shape[-2:]
        keep = torch.tril(torch.ones(tq, tk, dtype=torch.bool))
        energies = torch.where(keep, energies, torch.full_like(energies, -INF))
    if keys_mask is not None:  # mask_energies (:69-83): e * m + (1 - m) * -1e9
        m = keys_mask.to(energies.dtype).unsqueeze(1).unsqueeze(1)
        energies = energies * m + (1.0 - m) * -INF
    weights = torch.softmax(energies, dim=-1)
    if drop_mask is not None:
        weights = weights * drop_mask
    context = (weights @ v).permute(0, 2, 1, 3).reshape(queries.shape[0], queries.shape[1], dim)
    if heads > 1:
        context = dense(context, "output_proj")
    return context, weights


def _scoped_ln(p: Params, scope: str, x):
    prefix = scope + "/LayerNorm/" if scope else "LayerNorm/"
    return layer_norm(x, p[prefix + "gamma"], p[prefix + "beta"])


def transformer_feedforward(p: Params, scope: str, x):
    """feedforward_sublayer (encoders/transformer.py:266-288)."""
    normalized = _scoped_ln(p, scope, x)
    hidden = torch.relu(normalized @ p[scope + "/hidden_state/kernel"] + p[scope + "/hidden_state/bias"])
    return hidden @ p[scope + "/output/kernel"] + p[scope + "/output/bias"] + x


def transformer_encoder(p: Params, prefix: str, inputs: torch.Tensor, mask: torch.Tensor, depth: int,
                        heads: int, use_positional_encoding: bool = True) -> Dict[str, torch.Tensor]:
    """TransformerEncoder.temporal_states / output (encoders/transformer.py:174-330), no dropout.
    `inputs` [B,T,D]: the embedded input sequence; `mask` [B,T]."""
    x = inputs
    if use_positional_encoding:
        x = x + position_signal(x.shape[-1], x.shape[1], x.dtype)
    for i in range(depth):
        scope = "{}/layer_{}".format(prefix, i)
        normalized = _scoped_ln(p, scope + "/self_attention", x)
        ctx, _ = multihead_attention(p, scope + "/self_attention", normalized, normalized, normalized,
                                     mask, heads)
        x = ctx + x
        x = transformer_feedforward(p, scope + "/feedforward", x)
    states = _scoped_ln(p, prefix, x)
    return {"states": states, "mask": mask, "output": states.sum(dim=1)}


class TransformerDecoderSpec:
    def __init__(self, prefix: str, depth: int, heads_self: int, heads_enc: int, max_len: int,
                 tie_embeddings: bool = True, supress_unk: bool = False) -> None:
        self.prefix, self.depth, self.heads_self, self.heads_enc = prefix, depth, heads_self, heads_enc
        self.max_len, self.tie_embeddings, self.supress_unk = max_len, tie_embeddings, supress_unk


def cross_attention(p: Params, scope: str, queries, enc_states: Sequence[torch.Tensor],
                    enc_masks: Sequence[torch.Tensor], heads: Sequence[int], strategy: str = "serial",
                    heads_hier: Optional[int] = None) -> torch.Tensor:
    """The encoder-attention sublayer of a Transformer decoder layer over one or more encoders
    (attention/transformer_cross_layer.py:12-263; dropout callbacks = identity).
        serial        one `single` after the other (own LayerNorm and residual each), scopes enc_<i>
        parallel      one shared LayerNorm of the queries, sum of the per-encoder contexts + residual
        flat          states and masks concatenated on the time axis, one `single` in `scope` itself
        hierarchical  as parallel, but the contexts are attended over (scope enc_hier: queries
                      [B*T,1,d], keys = values = the stacked contexts [B*T,n,d], all-ones mask)"""
    def single(sc, q, states, mask, n_heads, normalize=True, residual=True):
        nq = _scoped_ln(p, sc, q) if normalize else q
        ctx, _ = multihead_attention(p, sc, nq, states, states, mask, n_heads)
        return ctx + q if residual else ctx

    if strategy == "serial":
        x = queries
        for i, (states, mask, n_heads) in enumerate(zip(enc_states, enc_masks, heads)):
            x = single("{}/enc_{}".format(scope, i), x, states, mask, n_heads)
        return x
    if strategy == "flat":
        return single(scope, queries, torch.cat(list(enc_states), 1), torch.cat(list(enc_masks), 1), heads[0])
    normalized = _scoped_ln(p, scope, queries)
    contexts = [single("{}/enc_{}".format(scope, i), normalized, states, mask, n_heads, False, False)
                for i, (states, mask, n_heads) in enumerate(zip(enc_states, enc_masks, heads))]
    if strategy == "parallel":
        return sum(contexts) + queries
    if strategy == "hierarchical":
        bsz, steps, dim = queries.shape
        stacked = torch.stack(contexts, 2).reshape(bsz * steps, len(contexts), dim)
        ones = torch.ones(bsz * steps, len(contexts), dtype=queries.dtype)
        ctx = single(scope + "/enc_hier", normalized.reshape(bsz * steps, 1, dim), stacked, ones, heads_hier,
                     False, False)
        return ctx.reshape(bsz, steps, dim) + queries
    raise ValueError(strategy)


def transformer_decoder_stack(p: Params, spec: TransformerDecoderSpec, inputs, mask, enc_states, enc_mask,
                              strategy: str = "serial", heads_enc: Optional[Sequence[int]] = None,
                              heads_hier: Optional[int] = None):
    """TransformerDecoder.layer (decoders/transformer.py:270-387).  `enc_states` / `enc_mask` are one
    tensor each (one encoder, the serial strategy) or lists (multi-source, any strategy)."""
    many = isinstance(enc_states, (list, tuple))
    states = list(enc_states) if many else [enc_states]
    masks = list(enc_mask) if many else [enc_mask]
    heads = list(heads_enc) if heads_enc is not None else [spec.heads_enc] * len(states)
    x = inputs
    for i in range(spec.depth):
        scope = "{}/layer_{}".format(spec.prefix, i)
        normalized = _scoped_ln(p, scope + "/self_attention", x)
        ctx, _ = multihead_attention(p, scope + "/self_attention", normalized, normalized, normalized,
                                     mask, spec.heads_self, masked=True)
        x = ctx + x
        x = cross_attention(p, scope + "/encdec_attention", x, states, masks, heads, strategy, heads_hier)
        x = transformer_feedforward(p, scope + "/feedforward", x)
    return _scoped_ln(p, spec.prefix, x)


def transformer_logits(p: Params, spec: TransformerDecoderSpec, states: torch.Tensor,
                       training: bool = False) -> torch.Tensor:
    """decoding_w / decoding_b (autoregressive.py:228-243) + supress_unk (:450-459).  The -1e9 <unk>
    column belongs to get_body's state_to_logits, i.e. to the run-time loops; the Transformer's training
    pass computes its logits itself (decoders/transformer.py:409-419) and never adds it (`training`)."""
    if spec.tie_embeddings:
        logits = states @ p[spec.prefix + "/word_embeddings"].t()
    else:
        logits = states @ p[spec.prefix + "/state_to_word_W"] + p[spec.prefix + "/state_to_word_b"]
    if spec.supress_unk and not training:
        pen = torch.zeros(logits.shape[-1], dtype=logits.dtype)
        pen[UNK] = -INF
        logits = logits + pen
    return logits


def transformer_decoder_train(p: Params, spec: TransformerDecoderSpec, enc: Dict[str, torch.Tensor],
                              tgt_ids: torch.Tensor, label_smoothing: Optional[float] = None) -> Dict[str, torch.Tensor]:
    """train_loop_result (decoders/transformer.py:389-447) + train_xents / train_loss
    (autoregressive.py:292-316).  tgt_ids [B,T] incl. </s>.  Inputs are embedded WITHOUT a
    position signal: the base-class `embed_input_symbols` is what the reference calls."""
    bsz, steps = tgt_ids.shape
    emb = p[spec.prefix + "/word_embeddings"]
    go = torch.full((bsz, 1), START, dtype=torch.int64)
    inputs = emb[torch.cat([go, tgt_ids[:, :-1]], dim=1)]
    mask = (tgt_ids != PAD).to(emb.dtype)
    states = transformer_decoder_stack(p, spec, inputs, mask, enc["states"], enc["mask"])
    logits = transformer_logits(p, spec, states, training=True)
    xent = sequence_xents(logits, tgt_ids, mask, label_smoothing)
    return {"states": states, "logits": logits, "xents": xent, "loss": xent.sum() / mask.sum(),
            "loss_sum": xent.sum(), "count": mask.sum()}


def transformer_decoder_greedy(p: Params, spec: TransformerDecoderSpec, enc: Dict[str, torch.Tensor]):
    """decoding_loop(train_mode=False) with TransformerDecoder.next_state (:485-518): the whole
    prefix is re-run every step; the mask column appended at a step is `not finished`."""
    bsz = enc["states"].shape[0]
    emb = p[spec.prefix + "/word_embeddings"]
    dim = emb.shape[1]
    prefix = {"seq": torch.zeros(bsz, 0, dim, dtype=emb.dtype), "mask": torch.zeros(bsz, 0, dtype=emb.dtype)}

    def next_output(embedded, finished):
        prefix["seq"] = torch.cat([prefix["seq"], embedded.unsqueeze(1)], 1)
        prefix["mask"] = torch.cat([prefix["mask"], (~finished).to(emb.dtype).unsqueeze(1)], 1)
        states = transformer_decoder_stack(p, spec, prefix["seq"], prefix["mask"], enc["states"], enc["mask"])
        return states[:, -1], None

    hist = autoregressive_loop(next_output, lambda o: transformer_logits(p, spec, o), lambda ids: emb[ids], bsz,
                               spec.max_len)
    return {"symbols": torch.stack(hist["symbols"]), "logits": torch.stack(hist["logits"]),
            "mask": torch.stack(hist["mask"])}


# ---------------------------------------------------------------------------
# ImageNet encoder (encoders/imagenet_encoder.py:131-240) over the slim VGG stack
# ---------------------------------------------------------------------------
VGG_BLOCKS = {"vgg_16": (2, 2, 3, 3, 3), "vgg_19": (2, 2, 4, 4, 4)}
VGG_CHANNELS = (64, 128, 256, 512, 512)


def vgg_features(p: Params, network_type: str, images: torch.Tensor,
                 spatial_layer: str) -> Dict[str, torch.Tensor]:
    """tensorflow/models research/slim nets/vgg.py (third-party, not vendored by the reference;
    restated from the published definition): repeat(conv 3x3 SAME + bias + relu), max_pool 2x2.
    images [B,H,W,3]; weights HWIO.  Returns the ImageNet part's tensors with
    `temporal_*` aliases so the RNN decoder oracle can attend over the flattened map
    (attention/base_attention.py:79-118)."""
    x = images.permute(0, 3, 1, 2)
    found = None
    for block, convs in enumerate(VGG_BLOCKS[network_type], 1):
        for i in range(1, convs + 1):
            name = "{}/conv{}/conv{}_{}".format(network_type, block, block, i)
            w = p[name + "/weights"].permute(3, 2, 0, 1)
            x = torch.relu(torch.nn.functional.conv2d(x, w, p[name + "/biases"], padding=1))
            if name == spatial_layer:
                found = x
                break
        if found is not None:
            break
        x = torch.nn.functional.max_pool2d(x, 2, 2)
        if "{}/pool{}".format(network_type, block) == spatial_layer:
            found = x
            break
    states = found.permute(0, 2, 3, 1).contiguous()
    bsz, h, w, c = states.shape
    return {"spatial_states": states, "spatial_mask": torch.ones(bsz, h, w, dtype=states.dtype),
            "output": states.mean(dim=(1, 2)),
            "temporal_states": states.reshape(bsz, h * w, c),
            "temporal_mask": torch.ones(bsz, h * w, dtype=states.dtype)}


# ---------------------------------------------------------------------------
# Beam search (decoders/beam_search_decoder.py:218-596)
# ---------------------------------------------------------------------------
def length_penalty(lengths: torch.Tensor, alpha: float) -> torch.Tensor:
    """_length_penalty (:560-573): ((5 + len) / 6) ** alpha on fp32 tensors.

    The fp32 division is IEEE-exact everywhere; the power is taken as the CORRECTLY ROUNDED
    fp32 power (evaluate in fp64, round once), which is what glibc's powf - behind Eigen's
    scalar pow in the TF-1.12 CPU kernel - delivers.  torch's vectorised fp32 pow differs
    from it in the last ulp for some arguments, so it is not used here."""
    base = (5.0 + lengths.to(torch.float32)) / 6.0
    return torch.pow(base.to(torch.float64), float(torch.tensor(alpha, dtype=torch.float32))
                     ).to(torch.float32)


def beam_step(logprobs: torch.Tensor, logprob_sum: torch.Tensor, lengths: torch.Tensor,
              finished: torch.Tensor, alpha: float):
    """Steps (1)-(8) of the beam body (:440-496) on fp32 tensors.

    logprobs [B,k,V] fp32, logprob_sum [B,k] fp32, lengths [B,k] int32, finished [B,k] bool."""
    bsz, k, vocab = logprobs.shape
    fmask = finished.to(torch.float32).unsqueeze(2)
    finished_row = torch.full((vocab,), -INF, dtype=torch.float32)
    finished_row[PAD] = 0.0
    lp = (1.0 - fmask) * logprobs + fmask * finished_row
    hyp_probs = logprob_sum.unsqueeze(2) + lp
    hyp_lengths = lengths + 1 - finished.to(torch.int32)
    scores = hyp_probs / length_penalty(hyp_lengths, alpha).unsqueeze(2)
    flat = scores.reshape(bsz, k * vocab)
    # tf.nn.top_k: descending, lower index first among equals -> stable sort on -score
    order = torch.sort(-flat, dim=1, stable=True).indices[:, :k]
    topk_scores = flat.gather(1, order)
    word_ids = (order % vocab).to(torch.int64)
    beam_ids = (order // vocab).to(torch.int32)
    next_lengths = hyp_lengths.gather(1, beam_ids.to(torch.int64))
    next_logprob_sum = hyp_probs.reshape(bsz, k * vocab).gather(1, order)
    next_finished = finished.gather(1, beam_ids.to(torch.int64)) | (word_ids == END)
    return topk_scores, word_ids, beam_ids, next_logprob_sum, next_lengths, next_finished


def beam_search(step_fn: Callable, init_state, first_logprobs: torch.Tensor, beam: int,
                max_steps: int, alpha: float, gather_state: Callable):
    """BeamSearchDecoder.outputs loop (:167-191,330-376) around a decoder body.

    first_logprobs [B*k, V]: log-softmax of the decoder body run once on the start symbol
    (get_initial_loop_state :218-328).  step_fn(state, word_ids[B*k]) -> (state, logprobs[B*k,V]).
    Returns token_ids [T,B,k] int64, scores [B,k], lengths, finished."""
    vocab = first_logprobs.shape[1]
    bsz = first_logprobs.shape[0] // beam
    logprob_sum = torch.full((bsz, beam), -INF, dtype=torch.float32)
    logprob_sum[:, 0] = 0.0
    lengths = torch.zeros(bsz, beam, dtype=torch.int32)
    finished = torch.zeros(bsz, beam, dtype=torch.bool)
    prev_logprobs = first_logprobs.reshape(bsz, beam, vocab).to(torch.float32)
    token_ids = torch.zeros(0, bsz, beam, dtype=torch.int64)
    scores = torch.zeros(bsz, beam, dtype=torch.float32)
    state = init_state
    step = 0
    # loop_continue_criterion (:330-355): step counts from 1 after the initial body run
    while step < max_steps and not bool(finished.all()):
        scores, words, beams, logprob_sum, lengths, finished = beam_step(
            prev_logprobs, logprob_sum, lengths, finished, alpha)
        flat_src = (torch.arange(bsz).unsqueeze(1) * beam + beams.to(torch.int64)).reshape(-1)
        state = gather_state(state, flat_src)
        token_ids = token_ids[:, torch.arange(bsz).unsqueeze(1), beams.to(torch.int64)]
        token_ids = torch.cat([token_ids, words.unsqueeze(0)], 0)
        state, lp = step_fn(state, words.reshape(-1), finished.reshape(-1))
        prev_logprobs = lp.reshape(bsz, beam, vocab).to(torch.float32)
        step += 1
    return {"token_ids": token_ids, "scores": scores, "lengths": lengths, "finished": finished}


# ---------------------------------------------------------------------------
# Trainer (trainers/generic_trainer.py:84-195) and tf.train.AdamOptimizer
# ---------------------------------------------------------------------------
def is_regularizable(name: str) -> bool:
    """BIAS_REGEX = r'[Bb]ias' (generic_trainer.py:17,87-91)."""
    import re
    return (not re.findall(r"[Bb]ias", name) and not name.startswith("vgg")
            and not name.startswith("Inception") and not name.startswith("resnet"))


def regularization(p: Params) -> Tuple[torch.Tensor, torch.Tensor]:
    reg = [v for n, v in p.items() if is_regularizable(n)]
    return sum(v.abs().sum() for v in reg), sum((v ** 2).sum() for v in reg)


def clip_by_norm(g: torch.Tensor, clip: float) -> torch.Tensor:
    norm = g.pow(2).sum().sqrt()
    return g * clip / torch.maximum(norm, torch.tensor(clip, dtype=g.dtype))


class AdamState:
    def __init__(self, p: Params) -> None:
        self.m = {n: torch.zeros_like(v) for n, v in p.items()}
        self.v = {n: torch.zeros_like(v) for n, v in p.items()}
        self.t = 0


def adam_step(p: Params, grads: Params, st: AdamState, lr: float = 1e-4, beta1: float = 0.9,
              beta2: float = 0.999, eps: float = 1e-8, clip_norm: Optional[float] = None) -> None:
    """apply_gradients with tf.train.AdamOptimizer defaults (generic_trainer.py:56-57,183-195)."""
    st.t += 1
    lr_t = lr * math.sqrt(1 - beta2 ** st.t) / (1 - beta1 ** st.t)
    with torch.no_grad():
        for n, v in p.items():
            g = grads[n]
            if clip_norm:
                g = clip_by_norm(g, clip_norm)
            st.m[n].mul_(beta1).add_(g, alpha=1 - beta1)
            st.v[n].mul_(beta2).addcmul_(g, g, value=1 - beta2)
            v.sub_(lr_t * st.m[n] / (st.v[n].sqrt() + eps))


def train_step(p: Params, spec: RNNDecoderSpec, enc_prefix: str, src_ids: torch.Tensor,
               train_inputs: torch.Tensor, st: AdamState, l1: float = 0.0, l2: float = 0.0,
               clip_norm: Optional[float] = None, lr: float = 1e-4) -> Dict[str, torch.Tensor]:
    """One CrossEntropyTrainer step on the Bahdanau model; returns the fetched losses."""
    for v in p.values():
        v.requires_grad_(True)
        v.grad = None
    enc = sentence_encoder(p, enc_prefix, src_ids)
    dec = decoder_train(p, spec, enc, train_inputs)
    l1_norm, l2_norm = regularization(p)
    total = dec["train_loss"] + l1 * l1_norm + l2 * l2_norm
    total.backward()
    grads = {n: (v.grad if v.grad is not None else torch.zeros_like(v)) for n, v in p.items()}
    for v in p.values():
        v.requires_grad_(False)
    adam_step(p, grads, st, lr=lr, clip_norm=clip_norm)
    return {"loss": dec["train_loss"].detach(), "l1": l1_norm.detach(), "l2": l2_norm.detach(),
            "grads": grads}


# ---------------------------------------------------------------------------
# parameter initialisation with the reference's initialisers
# ---------------------------------------------------------------------------
def orthogonal(rows: int, cols: int, gen: torch.Generator, dtype) -> torch.Tensor:
    """tf.orthogonal_initializer on a [rows, cols] kernel."""
    a = torch.randn(max(rows, cols), min(rows, cols), generator=gen, dtype=torch.float64)
    q, r = torch.linalg.qr(a)
    q = q * torch.sign(torch.diagonal(r))
    if rows < cols:
        q = q.t()
    return q[:rows, :cols].to(dtype).contiguous()


def init_bahdanau_params(vs: int, vt: int, es: int, he: int, et: int, hd: int, att: Optional[int],
                         out: int, maxout: bool, seed: int = 2574600, dtype=torch.float32,
                         enc_prefix: str = "sentence_encoder", att_prefix: str = "attention",
                         dec_prefix: str = "decoder") -> Params:
    gen = torch.Generator().manual_seed(seed)
    ctx = 2 * he
    att = att if att is not None else ctx

    def normal(*shape):
        return (torch.randn(*shape, generator=gen, dtype=torch.float64) * 0.001).to(dtype)

    p = {}
    p[enc_prefix + "_input/embedding_matrix_0"] = normal(vs, es)
    for d in ("fw", "bw"):
        cell = "{}/rnn_0_bidirectional/bidirectional_rnn/{}/OrthoGRUCell/".format(enc_prefix, d)
        p[cell + "gates/kernel"] = orthogonal(es + he, 2 * he, gen, dtype)
        p[cell + "gates/bias"] = torch.ones(2 * he, dtype=dtype)
        p[cell + "candidate/kernel"] = orthogonal(es + he, he, gen, dtype)
        p[cell + "candidate/bias"] = torch.zeros(he, dtype=dtype)
    p[enc_prefix + "/LayerNorm/gamma"] = torch.ones(ctx, dtype=dtype)
    p[enc_prefix + "/LayerNorm/beta"] = torch.zeros(ctx, dtype=dtype)
    p[att_prefix + "/Attention/attn_query_projection"] = normal(hd, att)
    p[att_prefix + "/attn_key_projection"] = normal(ctx, att)
    p[att_prefix + "/attn_similarity_v"] = normal(att)
    p[att_prefix + "/attn_projection_bias"] = torch.zeros(att, dtype=dtype)
    p[att_prefix + "/attn_bias"] = torch.zeros(1, dtype=dtype)
    p[dec_prefix + "/word_embeddings"] = normal(vt, et)
    p[dec_prefix + "/state_to_word_W"] = (
        (torch.rand(out, vt, generator=gen, dtype=torch.float64) - 0.5).to(dtype))
    p[dec_prefix + "/state_to_word_b"] = torch.zeros(vt, dtype=dtype)
    p[dec_prefix + "/initial_state/encoders_projection/kernel"] = normal(ctx, hd)
    p[dec_prefix + "/initial_state/encoders_projection/bias"] = torch.zeros(hd, dtype=dtype)
    cell = dec_prefix + "/attention_decoder/OrthoGRUCell/"
    p[cell + "gates/kernel"] = orthogonal(et + hd, 2 * hd, gen, dtype)
    p[cell + "gates/bias"] = torch.ones(2 * hd, dtype=dtype)
    p[cell + "candidate/kernel"] = orthogonal(et + hd, hd, gen, dtype)
    p[cell + "candidate/bias"] = torch.zeros(hd, dtype=dtype)
    cat = hd + et + ctx
    if maxout:
        pre = dec_prefix + "/attention_decoder/MaxoutProjection/MaxoutProjection/"
        p[pre + "kernel"] = normal(cat, 2 * out)
        p[pre + "bias"] = torch.zeros(2 * out, dtype=dtype)
    else:
        pre = dec_prefix + "/attention_decoder/dense/"
        p[pre + "kernel"] = normal(cat, out)
        p[pre + "bias"] = torch.zeros(out, dtype=dtype)
    return p


def randomize(p: Params, scale: float = 0.3, seed: int = 7) -> Params:
    """Replace the near-zero reference initialisation with O(scale) values so that parity tests
    exercise every non-linearity (N(0, 0.001^2) weights make tanh/softmax nearly linear)."""
    gen = torch.Generator().manual_seed(seed)
    out = {}
    for n, v in p.items():
        r = torch.randn(v.shape, generator=gen, dtype=torch.float64) * scale
        if n.endswith("gamma"):
            r = 1.0 + r
        out[n] = r.to(v.dtype)
    return out
