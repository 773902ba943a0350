"""K11 beam step: bit-exact integer bookkeeping against the fp32 CPU oracle."""
import pytest
import torch

from oracle import nm_oracle as O

pytestmark = pytest.mark.gpu


def _case(bsz, k, vocab, seed, finished_frac=0.3, ties=False):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(bsz, k, vocab, generator=g) * 3
    if ties:  # exact score ties inside a beam and across beams
        logits = torch.round(logits)
    logprobs = torch.log_softmax(logits, -1)
    logprob_sum = -torch.rand(bsz, k, generator=g) * 10
    if ties:
        logprob_sum = torch.round(logprob_sum)
    lengths = torch.randint(0, 20, (bsz, k), generator=g, dtype=torch.int32)
    finished = torch.rand(bsz, k, generator=g) < finished_frac
    if ties and k > 1:  # beams 0 and 1 are exact copies: every candidate ties across beams
        logprob_sum[:, 1] = logprob_sum[:, 0]
        logprobs[:, 1] = logprobs[:, 0]
        lengths[:, 1] = lengths[:, 0]
        finished[:, :2] = False
    return logprobs, logprob_sum, lengths, finished


@pytest.mark.parametrize("shape", [(1, 3, 70), (4, 8, 1000), (2, 8, 32000), (64, 8, 4099), (3, 1, 17)])
@pytest.mark.parametrize("alpha", [0.0, 0.6, 1.0])
@pytest.mark.parametrize("ties", [False, True])
def test_beam_step_bit_exact(shape, alpha, ties):
    from neuralmonkey_b200 import ops
    bsz, k, vocab = shape
    lp, ls, ln, fin = _case(bsz, k, vocab, seed=bsz * 131 + k, ties=ties)
    want = O.beam_step(lp, ls, ln, fin, alpha)
    got = ops.beam_step(lp.cuda(), ls.cuda(), ln.cuda(), fin.to(torch.uint8).cuda(), alpha)
    names = ("scores", "word_ids", "beam_ids", "logprob_sum", "lengths", "finished")
    for name, g_, w_ in zip(names, got, want):
        g_ = g_.cpu()
        if name == "finished":
            g_ = g_.bool()
        assert g_.dtype == w_.dtype or name == "finished", (name, g_.dtype, w_.dtype)
        assert torch.equal(g_, w_), name


def test_beam_step_initial_state():
    """logprob_sum = [0, -1e9, ...]: only beam 0 can be selected (beam_search_decoder.py:283-295)."""
    from neuralmonkey_b200 import ops
    bsz, k, vocab = 3, 8, 500
    g = torch.Generator().manual_seed(1)
    lp = torch.log_softmax(torch.randn(bsz, k, vocab, generator=g), -1)
    ls = torch.full((bsz, k), -O.INF)
    ls[:, 0] = 0.0
    ln = torch.zeros(bsz, k, dtype=torch.int32)
    fin = torch.zeros(bsz, k, dtype=torch.bool)
    want = O.beam_step(lp, ls, ln, fin, 0.6)
    got = ops.beam_step(lp.cuda(), ls.cuda(), ln.cuda(), fin.to(torch.uint8).cuda(), 0.6)
    assert torch.equal(got[2].cpu(), want[2]) and int(got[2].max()) == 0
    assert torch.equal(got[1].cpu(), want[1])


def test_beam_gather_rows():
    from neuralmonkey_b200 import ops
    bsz, k = 3, 4
    x = torch.arange(bsz * k * 10, dtype=torch.float32).view(bsz * k, 10)
    beams = torch.tensor([[3, 3, 0, 1], [0, 1, 2, 3], [2, 0, 0, 2]], dtype=torch.int32)
    out = ops.beam_gather(x.cuda(), beams.cuda(), bsz, k).cpu()
    idx = (torch.arange(bsz).unsqueeze(1) * k + beams.long()).reshape(-1)
    assert torch.equal(out, x[idx])
    ids = torch.arange(bsz * k, dtype=torch.int64).view(bsz * k, 1)
    out2 = ops.beam_gather(ids.cuda(), beams.cuda(), bsz, k).cpu()
    assert torch.equal(out2, ids[idx])
