"""CUDA-graph beam search over a TransformerDecoder with static-shape state.

A beam step launches ~150 small kernels (beam kernel, gathers, 6 layers x {LayerNorm, 4 small
GEMMs, attention core, FFN}, vocabulary projection, log-softmax); issued from Python one by one
the GPU idles between them and a step costs ~3.6 ms whatever the batch.  Here every per-hypothesis
tensor is a full-length buffer ([rows, max_time, ...]; keys beyond the current position are masked,
which gives them probability exactly 0), the position is a device scalar, and ONE captured graph
is replayed per step.  The host reads a "all hypotheses finished" flag every few steps; steps
run past that point only append <pad> columns (scores, lengths and order are idempotent once
everything is finished - beam_search_decoder.py:440-496), and the token history is trimmed to
the step the reference loop would have stopped at (:330-355).
"""
from typing import Any, Dict, List, Tuple

import torch

from neuralmonkey_b200 import ops, runtime


class TransformerBeamGraph:
    CHECK_EVERY = 8

    def __init__(self, bs_decoder, bsz: int, src_shapes: Tuple) -> None:
        self.bs = bs_decoder
        self.dec = bs_decoder.parent_decoder
        self.bsz, self.k = bsz, bs_decoder.beam_size
        self.rows = bsz * self.k
        self.max_steps = bs_decoder.max_steps
        self.tmax = self.max_steps + 1
        self.src_shapes = src_shapes
        dev = runtime.device()
        dec, rows, tmax = self.dec, self.rows, self.tmax
        dim, vocab = dec.dimension, len(dec.vocabulary)
        self.kv = [torch.zeros(rows, tmax, dim, device=dev) for _ in range(2 * dec.depth)]
        self.mask = torch.zeros(rows, tmax, device=dev)
        self.history = torch.zeros(rows, tmax, dtype=torch.int64, device=dev)
        self.pos = torch.zeros(1, dtype=torch.int64, device=dev)
        self.x = torch.zeros(rows, dim, device=dev)
        self.fin_rows = torch.zeros(rows, dtype=torch.bool, device=dev)
        self.logprobs = torch.zeros(bsz, self.k, vocab, device=dev)
        self.logprob_sum = torch.zeros(bsz, self.k, device=dev)
        self.lengths = torch.zeros(bsz, self.k, dtype=torch.int32, device=dev)
        self.finished = torch.zeros(bsz, self.k, dtype=torch.uint8, device=dev)
        self.scores = torch.zeros(bsz, self.k, device=dev)
        self.done = torch.zeros(tmax, dtype=torch.uint8, device=dev)
        self.enc_states = [torch.zeros(s, device=dev) for s in src_shapes]      # beam-tiled
        self.enc_masks = [torch.zeros(s[:2], device=dev) for s in src_shapes]
        self.graph = None
        self.cross = {}

    # -- the two halves of a step, written against the static buffers only -------------------
    def _decoder_part(self) -> None:
        dec = self.dec
        self.mask.index_copy_(1, self.pos, (~self.fin_rows).to(torch.float32).view(-1, 1))
        state, _ = dec._step_cached(self.x.unsqueeze(1), self.mask, self.kv, static_pos=self.pos)
        logits, lse, _argmax = dec.state_to_logits(state)
        self.logprobs.copy_(ops.log_softmax_from_lse(logits, lse).view(self.logprobs.shape))

    def _beam_part(self) -> None:
        bsz, k = self.bsz, self.k
        scores, words, beams, lsum, lens, fin = ops.beam_step(
            self.logprobs, self.logprob_sum, self.lengths, self.finished, self.bs.length_normalization)
        self.scores.copy_(scores)
        self.logprob_sum.copy_(lsum)
        self.lengths.copy_(lens)
        self.finished.copy_(fin)
        for buf in self.kv + [self.mask, self.history]:
            buf.copy_(ops.beam_gather(buf, beams, bsz, k))
        self.pos.add_(1)
        self.history.index_copy_(1, self.pos, words.view(-1, 1))
        self.x.copy_(self.dec.embed_input_symbols(words.view(-1)))
        self.fin_rows.copy_(fin.view(-1).to(torch.bool))
        self.done.index_copy_(0, self.pos, fin.min().view(1))

    def _step(self) -> None:
        self._beam_part()
        self._decoder_part()

    # -- one decode -------------------------------------------------------------------------------
    def _cross_projections(self) -> None:
        """Projected encoder keys / values into buffers whose addresses the graph can rely on."""
        dec = self.dec
        cache = dec.__dict__.setdefault("_batch_cache", {})
        for layer in range(dec.depth):
            for j, states in enumerate(self.enc_states):
                key = ("cross_kv", layer, j, states.data_ptr(), tuple(states.shape))
                cache.pop(key, None)
                ek, ev = dec._cross_kv(layer, j, states)          # fresh tensors for this batch
                if key not in self.cross:
                    self.cross[key] = (torch.empty_like(ek), torch.empty_like(ev))
                self.cross[key][0].copy_(ek)
                self.cross[key][1].copy_(ev)
                cache[key] = self.cross[key]

    def _init_state(self) -> None:
        dec = self.dec
        for buf in self.kv + [self.mask]:
            buf.zero_()
        self.history.zero_()
        self.done.zero_()
        self.pos.zero_()
        self.fin_rows.zero_()
        self.finished.zero_()
        self.lengths.zero_()
        self.scores.zero_()
        self.logprob_sum.fill_(-1e9)
        self.logprob_sum[:, 0] = 0.0
        go = torch.full((self.rows,), 1, dtype=torch.int64, device=self.x.device)   # <s>
        self.x.copy_(dec.embed_input_symbols(go))
        self._decoder_part()          # position 0: scores the first token (get_initial_loop_state)

    def run(self, tiled_states: List[torch.Tensor], tiled_masks: List[torch.Tensor]) -> Dict[str, Any]:
        dec, bsz, k = self.dec, self.bsz, self.k
        for dst, src in zip(self.enc_states, tiled_states):
            dst.copy_(src)
        for dst, src in zip(self.enc_masks, tiled_masks):
            dst.copy_(src if src is not None else torch.ones_like(dst))
        # the parent reads its encoders through these hooks; the static copies keep the addresses
        # the captured graph saw
        dec.encoder_states = lambda: self.enc_states
        dec.encoder_masks = lambda: self.enc_masks
        self._cross_projections()
        self._init_state()
        if self.graph is None:
            self._step()                       # warm-up outside the capture (lazy one-time setup)
            torch.cuda.synchronize()
            self._init_state()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self._step()
            # a capture records, it does not execute: the state is still the one after position 0
        steps = 0
        while steps < self.max_steps:
            chunk = min(self.CHECK_EVERY, self.max_steps - steps)
            for _ in range(chunk):
                self.graph.replay()
            steps += chunk
            done = self.done[1:steps + 1].cpu()
            if bool(done.any()):
                steps = int(torch.nonzero(done)[0]) + 1   # the step the reference loop stops after
                break
        token_ids = self.history[:, :steps + 1].view(bsz, k, steps + 1).permute(2, 0, 1).contiguous()
        return {"scores": self.scores.clone(), "token_ids": token_ids,
                "logprob_sum": self.logprob_sum.clone(), "lengths": self.lengths.clone(),
                "finished": self.finished.to(torch.bool), "logprobs": self.logprobs, "steps": steps}
