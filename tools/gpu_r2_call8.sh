#!/bin/bash
# Round 2, GPU call 8: CTA-pair (cta_group::2) GEMM instances - parity, then the shape sweep with and without pairs;
# launch list of a Transformer step; ncu evidence of the round (CSV pages only)
set -u
out=gpurun_out/call8
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout -k 10 600 python -m pytest tests/test_gpu_gemm_pair.py -m gpu -q -x > "$out/pair_tests.log" 2>&1
echo "pair tests exit $?" | tee -a "$out/summary.txt"
tail -n 15 "$out/pair_tests.log"
timeout -k 10 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_xent16.py -m gpu -q > "$out/gemm_tests.log" 2>&1
echo "gemm tests exit $?" | tee -a "$out/summary.txt"
tail -n 4 "$out/gemm_tests.log"
for cfg in "" "NMB200_TC_PAIR=1" "NMB200_TC_BN=256" "NMB200_TC_BN=256 NMB200_TC_PAIR=1" "NMB200_TC_BN=128 NMB200_TC_PAIR=1"; do
  echo "== sweep [$cfg]" | tee -a "$out/sweep.txt"
  env $cfg timeout -k 10 300 python tools/gemm_sweep.py --reps 40 >> "$out/sweep.txt" 2>> "$out/sweep.err"
  echo "sweep [$cfg] exit $?" | tee -a "$out/summary.txt"
done
cat "$out/sweep.txt"
# vocabulary path with pairs: the bench breakdown
NMB200_TC_PAIR=1 timeout -k 10 600 python bench.py --no-cpu-baseline --no-extras --breakdown > "$out/bench_pair.json" 2> "$out/bench_pair.log"
echo "bench pair exit $?" | tee -a "$out/summary.txt"
timeout -k 10 600 python bench.py --no-cpu-baseline --no-extras --breakdown > "$out/bench.json" 2> "$out/bench.log"
echo "bench exit $?" | tee -a "$out/summary.txt"
python - <<'PY'
import json
for f in ("bench_pair.json", "bench.json"):
    try:
        d = [json.loads(l) for l in open("gpurun_out/call8/" + f) if l.startswith("{")][-1]
        print(f, d["ms_per_step"], "ms", {k: round(v, 3) for k, v in list(d.get("breakdown_ms_per_step", {}).items())[:14]})
    except Exception as e:
        print(f, "unreadable", e)
PY
# launch list of a Transformer training step
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file "$out/transformer_launches.csv" \
    python bench_workloads.py transformer --no-cpu > "$out/transformer_list.log" 2>&1
echo "transformer list exit $?" | tee -a "$out/summary.txt"
timeout -k 10 1500 bash tools/gpu_ncu_r02.sh > "$out/ncu.log" 2>&1
echo "ncu exit $?" | tee -a "$out/summary.txt"
tail -n 12 "$out/ncu.log"
du -sh gpurun_out
