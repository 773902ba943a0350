#!/bin/bash
# Round 2, GPU call 12: decoder step kernel after the rows-per-cluster refactor (default 8 rows; 16 rows opt-in)
set -u
out=gpurun_out/call12
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout -k 10 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_beam.py -m gpu -q > "$out/decode_tests.log" 2>&1
echo "decode tests exit $?" | tee -a "$out/summary.txt"
tail -n 8 "$out/decode_tests.log" | cut -c1-220
timeout -k 10 300 python __graft_entry__.py smoke > "$out/smoke.log" 2>&1
echo "smoke exit $?" | tee -a "$out/summary.txt"
tail -n 2 "$out/smoke.log" | cut -c1-200
for rows in 8 16; do
  NMB200_DECSTEP_ROWS=$rows timeout -k 10 300 python tools/decstep_phases.py > "$out/phases_rows$rows.txt" 2> "$out/phases_rows$rows.log"
  echo "phases ($rows rows per cluster) exit $?" | tee -a "$out/summary.txt"
  cat "$out/phases_rows$rows.txt"
  NMB200_DECSTEP_ROWS=$rows timeout -k 10 600 python bench_workloads.py rnn_decode --no-cpu > "$out/rnn_decode_rows$rows.json" 2> "$out/rnn_decode_rows$rows.log"
  echo "rnn_decode ($rows rows per cluster) exit $?" | tee -a "$out/summary.txt"
  python tools/print_rnn_decode.py "$out/rnn_decode_rows$rows.json"
done
NMB200_DECSTEP_ROWS=16 timeout -k 10 600 python -m pytest tests/test_gpu_decode.py -m gpu -q > "$out/decode_tests_rows16.log" 2>&1
echo "decode tests (16 rows default) exit $?" | tee -a "$out/summary.txt"
tail -n 4 "$out/decode_tests_rows16.log" | cut -c1-220
du -sh gpurun_out
# the bench line once more: its roofline object now comes from per-call times taken on one stream
timeout -k 10 1200 python bench.py > "$out/bench.json" 2> "$out/bench.log"
echo "bench exit $?" | tee -a "$out/summary.txt"
python - <<'PY'
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/call12/bench.json") if l.startswith("{")][-1]
    print("en-de", d["ms_per_step"], "ms", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", d["roofline"]["frac"], "f16", d["roofline"]["f16_instances"]["frac"])
except Exception as e:
    print("bench unreadable", e)
PY
