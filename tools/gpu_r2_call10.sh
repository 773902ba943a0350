#!/bin/bash
# Round 2, GPU call 10: attention on the tensor cores (batched tcgen05 products), programmatic dependent launch of the
# GEMM kernels, Transformer step with both
set -u
out=gpurun_out/call10
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout -k 10 600 python -m pytest tests/test_gpu_mha_tc.py -m gpu -q -x > "$out/mha_tc.log" 2>&1
echo "mha tc tests exit $?" | tee -a "$out/summary.txt"
tail -n 30 "$out/mha_tc.log" | cut -c1-220
timeout -k 10 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_gemm_pair.py tests/test_gpu_ops.py -m gpu -q > "$out/gemm_ops.log" 2>&1
echo "gemm/ops tests exit $?" | tee -a "$out/summary.txt"
tail -n 5 "$out/gemm_ops.log" | cut -c1-220
NMB200_TC_PDL=1 timeout -k 10 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_xent16.py tests/test_gpu_perf_parity.py -m gpu -q > "$out/pdl_tests.log" 2>&1
echo "pdl tests exit $?" | tee -a "$out/summary.txt"
tail -n 5 "$out/pdl_tests.log" | cut -c1-220
for cfg in "" "NMB200_TC_PDL=1"; do
  echo "== sweep [$cfg]" | tee -a "$out/sweep.txt"
  env $cfg timeout -k 10 300 python tools/gemm_sweep.py --reps 40 --set transformer >> "$out/sweep.txt" 2>> "$out/sweep.err"
  echo "sweep [$cfg] exit $?" | tee -a "$out/summary.txt"
done
cat "$out/sweep.txt"
for cfg in "" "NMB200_MHA_TC=1" "NMB200_MHA_TC=1 NMB200_TC_PDL=1"; do
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
NMB200_MHA_TC=1 timeout -k 10 900 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_reference_inis.py -m gpu -q > "$out/transformer_tests.log" 2>&1
echo "transformer tests (tc attention) exit $?" | tee -a "$out/summary.txt"
tail -n 6 "$out/transformer_tests.log" | cut -c1-220
NMB200_MHA_TC=1 NMB200_TC_PDL=1 timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file "$out/transformer_launches.csv" python bench_workloads.py transformer --no-cpu > "$out/transformer_list.log" 2>&1
echo "transformer list exit $?" | tee -a "$out/summary.txt"
NMB200_TC_PDL=1 timeout -k 10 600 python bench.py --no-cpu-baseline --no-extras > "$out/bench_pdl.json" 2> "$out/bench_pdl.log"
echo "bench pdl exit $?" | tee -a "$out/summary.txt"
python - <<'PY'
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/call10/bench_pdl.json") if l.startswith("{")][-1]
    print("en-de with PDL:", d["ms_per_step"], "ms", round(d["value"]))
except Exception as e:
    print("bench pdl unreadable", e)
PY
du -sh gpurun_out
# decoder step with the weights staged by 2-D TMA
out=gpurun_out/call10
timeout -k 10 600 python -m pytest tests/test_gpu_decode.py -m gpu -q -x -k "step_kernel" > "$out/decstep_tests.log" 2>&1
echo "decoder step tests (both weight paths) exit $?" | tee -a "$out/summary.txt"
tail -n 12 "$out/decstep_tests.log" | cut -c1-220
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
    d = [json.loads(l) for l in open("gpurun_out/call10/rnn_decode_wtma.json") if l.startswith("{")][-1]
    for k in ("greedy", "beam8_batch", "beam8_latency"):
        e = d[k]; print(k, round(e["us_per_step"], 1), "us/step", round(e["tokens_per_s"]), "tok/s", e.get("step_breakdown_us"))
except Exception as e:
    print("rnn_decode unreadable", e)
PY
du -sh gpurun_out
