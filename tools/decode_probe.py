"""Timing probe of RNN decoding at the en-de perf shape: fused engine vs the step-by-step path.
   python tools/decode_probe.py [--no-stepwise]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from tests.helpers import build_bahdanau, feed   # noqa: E402

DIMS = dict(vs=32000, vt=32000, es=300, he=300, et=300, hd=300, out=300, maxout=False, max_len=50,
            supress_unk=False)


def ev():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = ev()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    from neuralmonkey_b200.decoders import BeamSearchDecoder
    model = build_bahdanau(**DIMS)
    dec = model["dec"]
    with torch.no_grad():
        dec.var("state_to_word_b")[2] = -1.0e4          # </s> never wins: full-length decodes
    out = {}
    g = torch.Generator().manual_seed(1)
    for bsz in (256, 64, 1):
        src = torch.randint(4, DIMS["vs"], (bsz, 50), generator=g)

        def greedy():
            feed(model, src, None, train=False)
            return dec.runtime_symbols

        dec.use_fused_decoding = True
        for _ in range(3):
            greedy()
        ms = timeit(greedy, 3)
        steps = int(dec.runtime_symbols.shape[0])
        out["greedy_fused_b{}".format(bsz)] = {"ms": ms, "steps": steps, "tokens_per_s": bsz * steps / ms * 1e3,
                                               "us_per_step": ms / steps * 1e3}
        if "--no-stepwise" not in sys.argv and bsz != 64:
            dec.use_fused_decoding = False
            greedy()
            ms = timeit(greedy, 1)
            out["greedy_stepwise_b{}".format(bsz)] = {"ms": ms, "tokens_per_s": bsz * steps / ms * 1e3,
                                                      "us_per_step": ms / steps * 1e3}
            dec.use_fused_decoding = True
    for bsz in (64, 1):
        src = torch.randint(4, DIMS["vs"], (bsz, 50), generator=g)
        bs = BeamSearchDecoder(name="bs{}".format(bsz), parent_decoder=dec, beam_size=8, max_steps=128,
                               length_normalization=0.6)

        def beam():
            feed(model, src, None, train=False)
            bs.reset_batch()
            return bs.outputs

        for fused in ((True, False) if "--no-stepwise" not in sys.argv else (True,)):
            bs.use_fused_step = fused
            for _ in range(3 if fused else 1):
                res = beam()
            ms = timeit(beam, 2 if fused else 1)
            steps = int(res.last_search_step_output.token_ids.shape[0] - 1)
            out["beam8_{}_b{}".format("fused" if fused else "stepwise", bsz)] = {
                "ms": ms, "steps": steps, "tokens_per_s": bsz * steps / ms * 1e3, "us_per_step": ms / steps * 1e3}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("wall", time.time() - t0, file=sys.stderr)
