#!/bin/bash
# Round 2, GPU call 7: L2 policy hints in the step kernel, threshold pre-filter in the beam top-k; ncu evidence of the round
set -u
out=gpurun_out/call7
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 600 python tools/decstep_phases.py > "$out/phases.txt" 2> "$out/phases.log"
echo "phases exit $?" | tee -a "$out/summary.txt"
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_beam.py tests/test_gpu_golden.py -m gpu -q > "$out/tests.log" 2>&1
echo "tests exit $?" | tee -a "$out/summary.txt"
timeout 600 python bench_workloads.py rnn_decode --no-cpu > "$out/rnn_decode.json" 2> "$out/rnn_decode.log"
echo "rnn_decode exit $?" | tee -a "$out/summary.txt"
cat "$out/phases.txt"; tail -n 4 "$out/tests.log"
python - <<'PY'
import json
d=json.load(open("gpurun_out/call7/rnn_decode.json"))
for k in ("greedy","beam8_batch","beam8_latency"):
    e=d[k]; print(k, round(e["us_per_step"],1), "us/step", round(e["tokens_per_s"]), "tok/s", e.get("step_breakdown_us"))
PY
timeout 1500 bash tools/gpu_ncu_r02.sh > "$out/ncu.log" 2>&1
echo "ncu exit $?" | tee -a "$out/summary.txt"
tail -n 8 "$out/ncu.log"
