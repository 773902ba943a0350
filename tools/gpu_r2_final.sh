#!/bin/bash
# Round 2, final 1-GPU call: what the driver runs at round end (GPU test suite, smoke, bench both arms) plus the ncu
# evidence of the final code
set -u
out=gpurun_out/final
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout -k 10 300 python __graft_entry__.py smoke > "$out/smoke.log" 2>&1
echo "smoke exit $?" | tee -a "$out/summary.txt"
tail -n 3 "$out/smoke.log" | cut -c1-200
timeout -k 10 1500 python -m pytest tests -m gpu -q > "$out/pytest_gpu.log" 2>&1
echo "pytest -m gpu exit $?" | tee -a "$out/summary.txt"
tail -n 8 "$out/pytest_gpu.log" | cut -c1-220
timeout -k 10 1200 python bench.py > "$out/bench.json" 2> "$out/bench.log"
echo "bench exit $?" | tee -a "$out/summary.txt"
timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 > "$out/bench_ref.json" 2> "$out/bench_ref.log"
echo "bench reference exit $?" | tee -a "$out/summary.txt"
python - <<'PY'
import json
for f in ("bench.json", "bench_ref.json"):
    try:
        d = [json.loads(l) for l in open("gpurun_out/final/" + f) if l.startswith("{")][-1]
        print(f, d.get("ms_per_step"), "ms", round(d["value"]), d["unit"], "e2e", d.get("e2e", {}).get("value"),
              "frac", (d.get("roofline") or {}).get("frac"), "parity", d.get("parity"))
        for k, v in (d.get("extra_workloads") or {}).items():
            print("   ", k, v.get("value") if isinstance(v, dict) else v, v.get("unit") if isinstance(v, dict) else "")
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file "$out/transformer_launches.csv" python bench_workloads.py transformer --no-cpu > "$out/transformer_list.log" 2>&1
echo "transformer list exit $?" | tee -a "$out/summary.txt"
timeout -k 10 1500 bash tools/gpu_ncu_r02.sh > "$out/ncu.log" 2>&1
echo "ncu exit $?" | tee -a "$out/summary.txt"
tail -n 6 "$out/ncu.log"
du -sh gpurun_out
