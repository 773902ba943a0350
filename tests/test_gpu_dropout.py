"""K15: dropout with the keep decisions drawn inside the kernel (csrc/dropout.cu).

The random stream cannot equal TensorFlow's, so the test pins what the reference's
`tf.nn.dropout` guarantees - each element kept with probability keep_prob and scaled by 1/keep_prob, the gradient
masked the same way - and what the step machinery needs: the same (seed, step, site) gives the same mask, a new
step or another call site a new one, also for a step replayed from a CUDA graph."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mask_of(y, x, keep):
    kept = y != 0
    assert torch.allclose(y[kept], x[kept] / keep, rtol=1e-6, atol=0)
    return kept


@pytest.mark.parametrize("keep", [0.9, 0.5, 0.1])
@pytest.mark.parametrize("n", [1 << 20, 1001, 3])
def test_kept_fraction_scale_and_gradient(keep, n):
    from neuralmonkey_b200 import ops, runtime
    runtime.advance_dropout()
    x = (torch.rand(n, device="cuda") + 0.5).requires_grad_(True)
    y = ops.dropout(x, keep)
    kept = _mask_of(y.detach(), x.detach(), keep)
    if n >= 1 << 20:
        assert abs(float(kept.float().mean()) - keep) < 4 * (keep * (1 - keep) / n) ** 0.5 + 1e-4
        # no structure along the quads a thread draws together
        lanes = kept.view(-1, 4).float().mean(0)
        assert float((lanes - keep).abs().max()) < 5e-3
    dy = torch.rand(n, device="cuda") + 0.5
    y.backward(dy)
    assert torch.equal(x.grad != 0, kept)
    assert torch.allclose(x.grad[kept], dy[kept] / keep, rtol=1e-6, atol=0)


def test_residual_rides_along_and_unaligned_views():
    from neuralmonkey_b200 import ops, runtime
    runtime.advance_dropout()
    base = torch.rand(4099, device="cuda") + 0.5
    x = base[1:4098].requires_grad_(True)                      # 4-byte aligned only
    res = (torch.rand(4097, device="cuda") + 3.0).requires_grad_(True)
    y = ops.dropout(x, 0.7, residual=res)
    plain = y.detach() - res.detach()
    kept = plain.abs() > 1e-3
    assert torch.allclose(plain[kept], x.detach()[kept] / 0.7, rtol=1e-5, atol=1e-6)
    assert 0.6 < float(kept.float().mean()) < 0.8
    y.sum().backward()
    assert torch.equal(res.grad, torch.ones_like(res))
    assert torch.equal(x.grad != 0, kept)


def test_same_counters_same_mask_new_step_new_mask():
    from neuralmonkey_b200 import lib, runtime
    state = runtime.dropout_state()
    runtime.advance_dropout()
    n = 1 << 16

    def mask(site):
        out = torch.empty(n, device="cuda")
        lib.call("nm_dropout_mask", lib.ptr(out), n, 0.5, lib.ptr(state), site, lib.stream())
        return out
    a, b, c = mask(7), mask(7), mask(8)
    assert torch.equal(a, b)
    assert 0.4 < float((a != c).float().mean()) < 0.6        # independent streams: half of the entries differ
    assert set(a.unique().tolist()) == {0.0, 2.0}
    runtime.advance_dropout()
    d = mask(7)
    assert 0.4 < float((a != d).float().mean()) < 0.6
    # the apply kernel draws the same decisions as the mask kernel
    x = torch.rand(n, device="cuda") + 0.5
    y = torch.empty_like(x)
    lib.call("nm_dropout_apply", lib.ptr(x), None, lib.ptr(y), n, 0.5, lib.ptr(state), 7, lib.stream())
    assert torch.equal(y != 0, d != 0)


def test_replayed_graph_draws_new_masks():
    from neuralmonkey_b200 import ops, runtime
    runtime.advance_dropout()
    x = torch.rand(1 << 14, device="cuda") + 0.5
    ops.dropout(x, 0.5)                                         # warm-up outside the capture
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y = ops.dropout(x, 0.5)
    masks = []
    for _ in range(3):
        runtime.advance_dropout()
        graph.replay()
        torch.cuda.synchronize()
        masks.append((y != 0).clone())
    assert 0.4 < float((masks[0] != masks[1]).float().mean()) < 0.6
    assert 0.4 < float((masks[1] != masks[2]).float().mean()) < 0.6


def test_nn_utils_paths(monkeypatch):
    """nn.utils.dropout: the kernel on CUDA tensors, a patched-in mask function still honoured."""
    from neuralmonkey_b200 import runtime
    from neuralmonkey_b200.nn import utils
    runtime.advance_dropout()
    x = torch.rand(64, 32, device="cuda") + 0.5
    assert utils.dropout(x, 1.0, True) is x and utils.dropout(x, 0.5, False) is x
    y = utils.dropout(x, 0.8, True)
    assert 0.7 < float((y != 0).float().mean()) < 0.9
    m = utils.dropout_mask((8, 4, 16, 16), 0.5, True, x.device)
    assert m.shape == (8, 4, 16, 16) and set(m.unique().tolist()) == {0.0, 2.0}

    def every_second(shape, keep_prob, train_mode, device):
        pattern = (torch.arange(shape[-1], device=device) % 2 == 0).float() / keep_prob
        return pattern.expand(tuple(shape)).clone()
    monkeypatch.setattr(utils, "dropout_mask", every_second)
    z = utils.dropout(x, 0.5, True, residual=x)
    assert torch.allclose(z[:, 0::2], 3.0 * x[:, 0::2]) and torch.equal(z[:, 1::2], x[:, 1::2])
