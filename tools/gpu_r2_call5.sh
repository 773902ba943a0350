#!/bin/bash
# Round 2, GPU call 5: warp-shuffle panel + division-free beam top-k + paired GRU launch: tests, decode numbers, train step, ncu
set -u
out=gpurun_out/call5
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_decode.py tests/test_gpu_beam.py tests/test_gpu_golden.py tests/test_gpu_model.py tests/test_gpu_ops.py tests/test_gpu_perf_parity.py tests/test_gpu_reference_inis.py -m gpu -q > "$out/tests.log" 2>&1
echo "tests exit $?" | tee -a "$out/summary.txt"
timeout 600 python bench_workloads.py rnn_decode --no-cpu > "$out/rnn_decode.json" 2> "$out/rnn_decode.log"
echo "rnn_decode exit $?" | tee -a "$out/summary.txt"
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --breakdown > "$out/bench.json" 2> "$out/bench.log"
echo "bench exit $?" | tee -a "$out/summary.txt"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_decoder_step|beam_local_topk" -s 90 -c 3 -o gpurun_out/call5/prof_decode -f \
    python tools/decode_probe.py --no-stepwise > "$out/ncu_decode.log" 2>&1
echo "ncu decode exit $?" | tee -a "$out/summary.txt"
tail -n 5 "$out/tests.log"
python - <<'PY'
import json
d=json.load(open("gpurun_out/call5/rnn_decode.json"))
for k in ("greedy","beam8_batch","beam8_latency"):
    e=d[k]; print(k, round(e["us_per_step"],1), "us/step", round(e["tokens_per_s"]), "tok/s", e.get("step_breakdown_us"))
b=json.load(open("gpurun_out/call5/bench.json")); print("train ms/step", b["ms_per_step"], b["value"], b["breakdown_ms_per_step"])
PY
