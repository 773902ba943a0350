#!/bin/bash
# 2-GPU call: the data-parallel equivalence test (eager with the early exchange and the range-wise update; whole step
# incl. NCCL in one graph; clean exit) and the two bench lines at N=2
set -u
out=gpurun_out/dp2
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -s > "$out/dp_tests.log" 2>&1
echo "dp tests exit $?" | tee -a "$out/summary.txt"
tail -n 12 "$out/dp_tests.log" | cut -c1-250
for wl in ende transformer; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 \
      bench.py --gpus 2 --steps 10 --warmup 3 --workload $wl > "$out/bench_${wl}_n2.json" 2> "$out/bench_${wl}_n2.log"
  echo "bench $wl N=2 exit $?" | tee -a "$out/summary.txt"
  python - "$out/bench_${wl}_n2.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("   ", d["ms_per_step"], "ms", round(d["value"]), d["unit"], "exposed", d.get("exposed_comm_ms_per_step"))
except Exception as e:
    print("    unreadable", e)
PY
done
