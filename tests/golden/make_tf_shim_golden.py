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
        return (rng.randn(*shape) * scale).astype(np.float32)

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
    np.savez_compressed(os.path.join(HERE, "tf_shim_golden.npz"), **out)
    print(sorted(out))


if __name__ == "__main__":
    main()
