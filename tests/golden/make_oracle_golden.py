"""Generate tests/golden/oracle_golden.npz: fixed-seed outputs of the CPU oracle at the shape
families of tests/bahdanau.ini and tests/transformer.ini / beamsearch.ini.

The reference itself cannot run here (TensorFlow 1.12), so these vectors
pin the ORACLE, not the reference: a CPU test checks that the oracle still reproduces them, and a
GPU test checks the CUDA path against the same numbers.  Run from the repo root:
    python tests/golden/make_oracle_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nm_oracle as O  # noqa: E402

BAHDANAU = dict(vs=60, vt=70, es=11, he=7, et=9, hd=8, att=None)   # tests/bahdanau.ini dims, maxout 9
TRANSFORMER = dict(vs=40, vt=44, dim=12, ff=20, depth=2, heads=3, max_len=7)


def batch(bsz, tx, ty, vs, vt, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(4, vs, (bsz, tx), generator=g)
    tgt = torch.randint(4, vt, (bsz, ty), generator=g)
    src_len = torch.randint(1, tx + 1, (bsz,), generator=g)
    tgt_len = torch.randint(1, ty, (bsz,), generator=g)
    src_len[0], tgt_len[0] = tx, ty - 1
    for b in range(bsz):
        src[b, src_len[b]:] = 0
        tgt[b, tgt_len[b]] = 2
        tgt[b, tgt_len[b] + 1:] = 0
    return src, tgt


def transformer_params(cfg, seed):
    g = torch.Generator().manual_seed(seed)
    d, f = cfg["dim"], cfg["ff"]
    p = {}

    def rnd(*shape):
        return torch.randn(*shape, generator=g) * 0.2

    def ln(pre):
        p[pre + "/LayerNorm/gamma"] = 1.0 + rnd(d)
        p[pre + "/LayerNorm/beta"] = rnd(d)

    def att(pre):
        for n in ("query_proj", "keys_proj", "vals_proj", "output_proj"):
            p["{}/{}/kernel".format(pre, n)] = rnd(d, d)

    def ff(pre):
        ln(pre)
        p[pre + "/hidden_state/kernel"], p[pre + "/hidden_state/bias"] = rnd(d, f), rnd(f)
        p[pre + "/output/kernel"], p[pre + "/output/bias"] = rnd(f, d), rnd(d)

    for i in range(cfg["depth"]):
        ln("encoder/layer_{}/self_attention".format(i))
        att("encoder/layer_{}/self_attention".format(i))
        ff("encoder/layer_{}/feedforward".format(i))
        ln("decoder/layer_{}/self_attention".format(i))
        att("decoder/layer_{}/self_attention".format(i))
        ln("decoder/layer_{}/encdec_attention/enc_0".format(i))
        att("decoder/layer_{}/encdec_attention/enc_0".format(i))
        ff("decoder/layer_{}/feedforward".format(i))
    ln("encoder")
    ln("decoder")
    p["input_sequence/embedding_matrix_0"] = rnd(cfg["vs"], d)
    p["decoder/word_embeddings"] = rnd(cfg["vt"], d)
    return p


def transformer_beam(p, spec, enc, beam, max_steps, alpha):
    emb = p["decoder/word_embeddings"]
    states = enc["states"].repeat_interleave(beam, 0)
    emask = enc["mask"].repeat_interleave(beam, 0)
    rows = states.shape[0]

    def run(seq, mask):
        out = O.transformer_decoder_stack(p, spec, seq, mask, states, emask)
        return torch.log_softmax(O.transformer_logits(p, spec, out[:, -1]), -1)

    seq0 = emb[torch.full((rows,), O.START, dtype=torch.int64)].unsqueeze(1)
    mask0 = torch.ones(rows, 1)

    def step_fn(state, words, finished):
        seq = torch.cat([state[0], emb[words].unsqueeze(1)], 1)
        mask = torch.cat([state[1], (~finished).to(emb.dtype).unsqueeze(1)], 1)
        return (seq, mask), run(seq, mask)

    return O.beam_search(step_fn, (seq0, mask0), run(seq0, mask0), beam, max_steps, alpha,
                         lambda st, idx: (st[0][idx], st[1][idx]))


def compute():
    out = {}
    # ---- Bahdanau toy -------------------------------------------------------------------
    c = BAHDANAU
    p = O.randomize(O.init_bahdanau_params(c["vs"], c["vt"], c["es"], c["he"], c["et"], c["hd"], c["att"],
                                           out=9, maxout=True), scale=0.3, seed=7)
    spec = O.RNNDecoderSpec("decoder", "attention", 10, "maxout", True)
    src, tgt = batch(5, 8, 7, c["vs"], c["vt"], seed=21)
    enc = O.sentence_encoder(p, "sentence_encoder", src)
    tr = O.decoder_train(p, spec, enc, tgt.t())
    gr = O.decoder_greedy(p, spec, enc, tgt.t())
    out.update({"b_src": src, "b_tgt": tgt, "b_enc_output": enc["output"], "b_train_loss": tr["train_loss"],
                "b_train_xents": tr["train_xents"], "b_greedy_symbols": gr["output_symbols"],
                "b_runtime_loss": gr["runtime_loss"]})
    # ---- Transformer toy ----------------------------------------------------------------
    t = TRANSFORMER
    tp = transformer_params(t, seed=5)
    tspec = O.TransformerDecoderSpec("decoder", t["depth"], t["heads"], t["heads"], t["max_len"], True, False)
    tsrc, ttgt = batch(4, 6, 6, t["vs"], t["vt"], seed=33)
    emb = tp["input_sequence/embedding_matrix_0"]
    mask = (tsrc != 0).float()
    tenc = O.transformer_encoder(tp, "encoder", emb[tsrc] * (mask * t["dim"] ** 0.5).unsqueeze(-1), mask,
                                 t["depth"], t["heads"])
    ttr = O.transformer_decoder_train(tp, tspec, tenc, ttgt)
    tgr = O.transformer_decoder_greedy(tp, tspec, tenc)
    tbm = transformer_beam(tp, tspec, tenc, 3, 6, 0.6)
    out.update({"t_src": tsrc, "t_tgt": ttgt, "t_enc_output": tenc["output"], "t_train_loss": ttr["loss"],
                "t_greedy_symbols": tgr["symbols"], "t_beam_tokens": tbm["token_ids"],
                "t_beam_scores": tbm["scores"], "t_beam_lengths": tbm["lengths"]})
    for k, v in tp.items():
        out["tp::" + k] = v
    for k, v in p.items():
        out["bp::" + k] = v
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()}


if __name__ == "__main__":
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"), **compute())
    print("written")
