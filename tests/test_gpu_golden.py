"""The CUDA path against the committed golden vectors (tests/golden/oracle_golden.npz): the same
numbers the CPU suite pins the oracle to, without running the oracle."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import build_bahdanau, feed
from tests.test_gpu_transformer import build_transformer, feed_transformer

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz")


def _params(golden, prefix):
    return {k[len(prefix):]: torch.from_numpy(golden[k]) for k in golden.files if k.startswith(prefix)}


def test_bahdanau_against_golden():
    from neuralmonkey_b200 import ops
    g = np.load(GOLDEN)
    try:
        ops.set_gemm_backend("simt")
        model = build_bahdanau(vs=60, vt=70, es=11, he=7, et=9, hd=8, out=9, maxout=True, max_len=10,
                               supress_unk=True)
        model["arena"].load_dict(_params(g, "bp::"))
        src, tgt = torch.from_numpy(g["b_src"]), torch.from_numpy(g["b_tgt"])
        feed(model, src, tgt, train=True)
        assert np.abs(model["enc"].output.detach().cpu().numpy() - g["b_enc_output"]).max() < 5e-5
        assert abs(float(model["dec"].train_loss) - float(g["b_train_loss"])) < 1e-4
        assert np.abs(model["dec"].train_xents.detach().cpu().numpy() - g["b_train_xents"]).max() < 2e-4
        feed(model, src, tgt, train=False)
        assert (model["dec"].runtime_symbols.cpu().numpy() == g["b_greedy_symbols"]).all()
        assert abs(float(model["dec"].runtime_loss) - float(g["b_runtime_loss"])) < 2e-4
    finally:
        ops.set_gemm_backend("auto")


@pytest.mark.parametrize("graph", [True, False])
def test_transformer_against_golden(graph):
    from neuralmonkey_b200 import ops
    from neuralmonkey_b200.decoders import BeamSearchDecoder
    g = np.load(GOLDEN)
    try:
        ops.set_gemm_backend("simt")
        model = build_transformer(vs=40, vt=44, dim=12, ff=20, depth=2, heads=3, max_len=7)
        model["arena"].load_dict(_params(g, "tp::"))
        src, tgt = torch.from_numpy(g["t_src"]), torch.from_numpy(g["t_tgt"])
        feed_transformer(model, src, tgt, train=True)
        assert np.abs(model["enc"].output.detach().cpu().numpy() - g["t_enc_output"]).max() < 2e-4
        assert abs(float(model["dec"].train_loss) - float(g["t_train_loss"])) < 1e-4
        feed_transformer(model, src, tgt, train=False)
        assert (model["dec"].runtime_symbols.cpu().numpy() == g["t_greedy_symbols"]).all()
        bs = BeamSearchDecoder(name="bs", parent_decoder=model["dec"], beam_size=3, max_steps=6,
                               length_normalization=0.6)
        bs.use_cuda_graph = graph
        bs.GRAPH_AFTER = 1        # capture on first use (default: second occurrence of a shape)
        feed_transformer(model, src, None, train=False)
        bs.reset_batch()
        bs.batch_size = src.shape[0]
        out = bs.outputs
        assert (out.last_search_step_output.token_ids.cpu().numpy()[1:] == g["t_beam_tokens"]).all()
        assert np.abs(out.last_search_step_output.scores.cpu().numpy() - g["t_beam_scores"]).max() < 2e-4
        assert (out.last_search_state.lengths.cpu().numpy() == g["t_beam_lengths"]).all()
    finally:
        ops.set_gemm_backend("auto")
