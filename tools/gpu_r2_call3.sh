#!/bin/bash
# Round 2, GPU call 3: decode step breakdown, perf-shape parity, the rebuilt bench (default run), re-run of the two files that had tolerance failures
set -u
out=gpurun_out/call3
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 600 python bench_workloads.py rnn_decode --no-cpu > "$out/rnn_decode.json" 2> "$out/rnn_decode.log"
echo "rnn_decode exit $?" | tee -a "$out/summary.txt"
timeout 900 python -m pytest tests/test_gpu_perf_parity.py tests/test_gpu_xent16.py tests/test_gpu_decode.py -m gpu -q > "$out/parity.log" 2>&1
echo "parity/xent16/decode exit $?" | tee -a "$out/summary.txt"
NMB200_UNVERIFIED=1 timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q > "$out/variants.log" 2>&1
echo "variants exit $?" | tee -a "$out/summary.txt"
timeout 1200 python bench.py --steps 10 --warmup 3 --breakdown > "$out/bench.json" 2> "$out/bench.log"
echo "bench exit $?" | tee -a "$out/summary.txt"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > "$out/bench_ref.json" 2> "$out/bench_ref.log"
echo "bench reference exit $?" | tee -a "$out/summary.txt"
for f in parity variants; do echo "== $f"; tail -n 6 "$out/$f.log"; done
cat "$out/rnn_decode.json"; echo; tail -n 40 "$out/bench.log" | grep "^#" | head -40
