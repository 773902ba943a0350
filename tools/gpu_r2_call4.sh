#!/bin/bash
# Round 2, GPU call 4: rewritten (small-code) decoder step + warp-level beam top-k: tests, numbers, ncu of the step kernel;
# the reference's INIs unchanged; variants without the switch
set -u
out=gpurun_out/call4
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_beam.py tests/test_gpu_golden.py tests/test_gpu_model.py -m gpu -q > "$out/decode_tests.log" 2>&1
echo "decode/beam/golden/model tests exit $?" | tee -a "$out/summary.txt"
timeout 600 python bench_workloads.py rnn_decode --no-cpu > "$out/rnn_decode.json" 2> "$out/rnn_decode.log"
echo "rnn_decode exit $?" | tee -a "$out/summary.txt"
timeout 900 python -m pytest tests/test_gpu_reference_inis.py tests/test_gpu_variants.py -m gpu -q > "$out/inis.log" 2>&1
echo "reference inis + variants exit $?" | tee -a "$out/summary.txt"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_decoder_step -s 60 -c 2 -o gpurun_out/call4/prof_decstep -f \
    python tools/decode_probe.py --no-stepwise > "$out/ncu_decstep.log" 2>&1
echo "ncu decstep exit $?" | tee -a "$out/summary.txt"
for f in decode_tests inis; do echo "== $f"; tail -n 6 "$out/$f.log"; done
python - <<'PY'
import json
d=json.load(open("gpurun_out/call4/rnn_decode.json"))
for k in ("greedy","beam8_batch","beam8_latency"):
    e=d[k]; print(k, round(e["us_per_step"],1), "us/step", round(e["tokens_per_s"]), "tok/s", e.get("step_breakdown_us"))
PY
