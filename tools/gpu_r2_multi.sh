#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus N): DP equivalence test, en-de and transformer bench at N ranks
set -u
N=${1:-2}
WLS=${2:-"ende transformer"}
out=gpurun_out/multi$N
mkdir -p "$out"
export PYTHONUNBUFFERED=1
if [ "$N" = "2" ]; then
  timeout 300 python -m pytest tests/test_gpu_mha_tc.py -m gpu -q > "$out/mha_tc.log" 2>&1
  echo "mha tc tests exit $?" | tee -a "$out/summary.txt"
  timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -s > "$out/dp_tests.log" 2>&1
  echo "dp tests exit $?" | tee -a "$out/summary.txt"
fi
for wl in $WLS; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
      bench.py --gpus $N --steps 10 --warmup 3 --workload $wl > "$out/bench_${wl}_n$N.json" 2> "$out/bench_${wl}_n$N.log"
  echo "bench $wl N=$N exit $?" | tee -a "$out/summary.txt"
done
# the same two workloads on one GPU of the same box, for the ratio (N = 2 only: a larger box is charged N times)
[ "$N" = "2" ] && for wl in ende transformer; do
  timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --workload $wl --no-extras --no-cpu-baseline > "$out/bench_${wl}_n1.json" 2> "$out/bench_${wl}_n1.log"
  echo "bench $wl N=1 exit $?" | tee -a "$out/summary.txt"
done
[ -f "$out/dp_tests.log" ] && tail -n 5 "$out/dp_tests.log"
python - "$out" "$N" <<'PY'
import json, sys
out, n = sys.argv[1], sys.argv[2]
for wl in ("ende", "transformer"):
    try:
        b = json.load(open("{}/bench_{}_n{}.json".format(out, wl, n)))
        print(wl, "N=" + n, round(b["ms_per_step"], 3), "ms", round(b["value"]), b["unit"], "exposed comm ms", b.get("exposed_comm_ms_per_step"))
        a = json.load(open("{}/bench_{}_n1.json".format(out, wl)))
        print("   N=1", round(a["ms_per_step"], 3), "ms", round(a["value"]), "| speed-up", round(b["value"] / a["value"], 3))
    except Exception as exc:
        print(wl, "failed:", exc)
PY
