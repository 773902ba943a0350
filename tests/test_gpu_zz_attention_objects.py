"""An RNN decoder with scaled-dot attention OBJECTS (attention.ScaledDotProdAttention / MultiHeadAttention,
tests/post-edit.ini's topology) on the GPU against the oracle.

Written after round 2's GPU budget was spent: the host composition is checked on the CPU over stand-in
operations (tests/test_host_model_cpu.py::test_rnn_decoder_with_scaled_dot_attention_objects) and every
operation it calls is GPU-verified in other compositions (the Transformer decoder's cross-attention makes the same
`ops.mha_core` call), but THIS test has never run.  It is therefore opt-in - `NMB200_RUN_UNRUN_GPU_TESTS=1` -
so that the suite the driver runs holds only tests that have been seen green on a B200; `bench.py` runs it in a
separate process and records the outcome under `extra_workloads.late_gpu_checks`."""
import os

import pytest
import torch

from oracle import nm_oracle as O
from tests.helpers import max_abs, oracle_params_for, random_batch
from tests.test_host_model_cpu import _post_edit_feed, _post_edit_model, _post_edit_oracle

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("NMB200_RUN_UNRUN_GPU_TESTS", "0") != "1",
                                 reason="never run on a GPU yet: opt in with NMB200_RUN_UNRUN_GPU_TESTS=1")]


@pytest.mark.parametrize("backend,tol", [("simt", 5e-5), ("auto", 1e-2)])
@pytest.mark.parametrize("heads", [3, 1])
def test_rnn_decoder_with_scaled_dot_attention_objects_on_the_gpu(heads, backend, tol):
    from neuralmonkey_b200 import ops
    try:
        ops.set_gemm_backend(backend)
        model = _post_edit_model(heads)
        arena, dec = model["arena"], model["dec"]
        params = oracle_params_for(model)
        arena.load_dict(params)
        src, tgt = random_batch(5, 8, 7, model["vs"], model["vt"], seed=4)
        trans, _ = random_batch(5, 8, 7, model["vt"], model["vt"], seed=5)
        _post_edit_feed(model, src, trans, tgt, train=True)
        p64 = {n: v.double().requires_grad_(True) for n, v in params.items()}
        spec = O.RNNDecoderSpec("decoder", None, 8, "tanh", False)
        enc, attend = _post_edit_oracle(p64, heads, src, trans)
        odec = O.decoder_train(p64, spec, enc, tgt.t(), attend=attend)
        assert max_abs(dec.train_output_states, odec["train_output_states"]) < tol
        assert abs(float(dec.train_loss) - float(odec["train_loss"])) < max(tol, 1e-5)
        arena.zero_grad()
        dec.train_loss.backward()
        odec["train_loss"].backward()
        gtol = 3e-4 if backend == "simt" else 2e-2
        for name, grad in arena.named_grads().items():
            want = p64[name].grad
            want = torch.zeros_like(p64[name]) if want is None else want
            err = float((grad.double().cpu() - want.reshape(grad.shape)).norm())
            assert err <= gtol * float(want.norm()) + 1e-6, (name, err, float(want.norm()))
        _post_edit_feed(model, src, trans, tgt, train=False)
        enc, attend = _post_edit_oracle(params, heads, src, trans)
        og = O.decoder_greedy(params, spec, enc, tgt.t(), attend=attend)
        assert dec.decode_engine is None
        assert max_abs(dec.runtime_logits, og["runtime_logits"]) < 10 * tol
        if backend == "simt":
            assert bool((dec.runtime_symbols.cpu() == og["output_symbols"]).all())
        _post_edit_feed(model, src, trans, tgt, train=True)
        out = model["trainer"].train_step()
        assert float(out["losses"][0]) > 0.0
    finally:
        ops.set_gemm_backend("auto")
