#!/bin/bash
# Round 2, GPU call 11: attention epilogues with staged (coalesced) loads / stores, weight gradients on a second
# stream, decoder step with TMA-staged weights
set -u
out=gpurun_out/call11
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout -k 10 600 python -m pytest tests/test_gpu_mha_tc.py -m gpu -q -x > "$out/mha_tc.log" 2>&1
echo "mha tc tests exit $?" | tee -a "$out/summary.txt"
tail -n 12 "$out/mha_tc.log" | cut -c1-220
timeout -k 10 600 python -m pytest tests/test_gpu_decode.py -m gpu -q -x -k "step_kernel" > "$out/decstep_tests.log" 2>&1
echo "decoder step tests (both weight paths) exit $?" | tee -a "$out/summary.txt"
tail -n 12 "$out/decstep_tests.log" | cut -c1-220
NMB200_WGRAD_STREAM=1 NMB200_MHA_TC=1 timeout -k 10 900 python -m pytest tests/test_gpu_perf_parity.py tests/test_gpu_model.py tests/test_gpu_transformer.py tests/test_gpu_reference_inis.py -m gpu -q > "$out/model_tests_side.log" 2>&1
echo "model tests (second stream + tc attention) exit $?" | tee -a "$out/summary.txt"
tail -n 6 "$out/model_tests_side.log" | cut -c1-220
for cfg in "" "NMB200_MHA_TC=1" "NMB200_WGRAD_STREAM=1" "NMB200_MHA_TC=1 NMB200_WGRAD_STREAM=1"; do
  tag=$(echo "$cfg" | tr -c 'A-Za-z0-9' '_')
  env $cfg timeout -k 10 600 python bench_workloads.py transformer --no-cpu > "$out/transformer_$tag.json" 2> "$out/transformer_$tag.log"
  echo "transformer [$cfg] exit $?" | tee -a "$out/summary.txt"
  python - "$out/transformer_$tag.json" <<'PY'
import json, sys
try:
    d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    print("   ", d["ms_per_step"], "ms/step", round(d["value"]), "tok/s", "loss", d.get("last_loss"))
except Exception as e:
    print("    unreadable", e)
PY
done
for cfg in "" "NMB200_WGRAD_STREAM=1"; do
  tag=$(echo "$cfg" | tr -c 'A-Za-z0-9' '_')
  env $cfg timeout -k 10 600 python bench.py --no-cpu-baseline --no-extras > "$out/bench_$tag.json" 2> "$out/bench_$tag.log"
  echo "bench [$cfg] exit $?" | tee -a "$out/summary.txt"
  python - "$out/bench_$tag.json" <<'PY'
import json, sys
try:
    d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    print("    en-de:", d["ms_per_step"], "ms", round(d["value"]), "tok/s e2e", round(d["e2e"]["value"]), "parity", d.get("parity"))
except Exception as e:
    print("    unreadable", e)
PY
done
NMB200_DECSTEP_WTMA=1 timeout -k 10 300 python tools/decstep_phases.py > "$out/phases_wtma.txt" 2> "$out/phases_wtma.log"
echo "phases (staged weights) exit $?" | tee -a "$out/summary.txt"
cat "$out/phases_wtma.txt"
NMB200_DECSTEP_WTMA=1 timeout -k 10 600 python -m pytest tests/test_gpu_decode.py -m gpu -q > "$out/decode_tests_wtma.log" 2>&1
echo "decode tests (staged weights) exit $?" | tee -a "$out/summary.txt"
tail -n 5 "$out/decode_tests_wtma.log" | cut -c1-220
NMB200_DECSTEP_WTMA=1 timeout -k 10 600 python bench_workloads.py rnn_decode --no-cpu > "$out/rnn_decode_wtma.json" 2> "$out/rnn_decode_wtma.log"
echo "rnn_decode (staged weights) exit $?" | tee -a "$out/summary.txt"
python - <<'PY'
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/call11/rnn_decode_wtma.json") if l.startswith("{")][-1]
    for k in ("greedy", "beam8_batch", "beam8_latency"):
        e = d[k]; print(k, round(e["us_per_step"], 1), "us/step", round(e["tokens_per_s"]), "tok/s", e.get("step_breakdown_us"))
except Exception as e:
    print("rnn_decode unreadable", e)
PY
NMB200_MHA_TC=1 NMB200_WGRAD_STREAM=1 timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file "$out/transformer_launches.csv" python bench_workloads.py transformer --no-cpu > "$out/transformer_list.log" 2>&1
echo "NMB200_MHA_TC=1 NMB200_WGRAD_STREAM=1" > "$out/transformer_env.txt"
echo "transformer list exit $?" | tee -a "$out/summary.txt"
du -sh gpurun_out
