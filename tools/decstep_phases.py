"""Cycle stamps of the decoder step kernel's phases (CTA 0) at the en-de shape.  python tools/decstep_phases.py"""
import sys
import torch
sys.path.insert(0, ".")
from tests.helpers import build_bahdanau, feed   # noqa: E402
from neuralmonkey_b200 import lib                 # noqa: E402

DIMS = dict(vs=32000, vt=32000, es=300, he=300, et=300, hd=300, out=300, maxout=False, max_len=50, supress_unk=False)
model = build_bahdanau(**DIMS)
dec = model["dec"]
with torch.no_grad():
    dec.var("state_to_word_b")[2] = -1.0e4
names = ["cluster up", "P1 gates", "P2 state", "P3 query", "P4 attention", "join", "P5 output"]
for bsz in (256, 8, 1):
    src = torch.randint(4, 32000, (bsz, 50))
    eng = dec.decode_engine
    eng.use_cuda_graph = False
    feed(model, src, None, train=False)
    _ = dec.runtime_symbols
    buf = torch.zeros(8, dtype=torch.int64, device="cuda")
    lib.call("nm_attn_decoder_step_debug", lib.ptr(buf))
    feed(model, src, None, train=False)
    _ = dec.runtime_symbols
    torch.cuda.synchronize()
    lib.call("nm_attn_decoder_step_debug", None)
    t = buf.cpu().tolist()
    d = [t[i + 1] - t[i] for i in range(7)]
    print("batch", bsz, "total cycles", t[7] - t[0], {n: v for n, v in zip(names, d)})
