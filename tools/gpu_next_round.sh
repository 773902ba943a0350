#!/bin/bash
# First multi-GPU call of the next round (gpurun --gpus 4 or 8):  bash tools/gpu_next_round.sh <N>
# Separates the two suspects DESIGN.md section 6 leaves open for the N >= 4 numbers of the range-wise optimizer update:
# the clock sampler that used to spawn nvidia-smi during the timed loop (now one NVML handle), and the 4-byte
# all-reduce of the token count inside the captured graph.  en-de only, both modes, host time per step printed.
set -u
N=${1:-4}
out=gpurun_out/next_round
mkdir -p "$out"
export PYTHONUNBUFFERED=1
for mode in 0 1; do
  NMB200_RANGE_UPDATE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 2955$mode bench.py --gpus $N --steps 20 --warmup 5 --workload ende \
      > "$out/bench_ende_n${N}_range$mode.json" 2> "$out/bench_ende_n${N}_range$mode.log"
  echo "bench ende N=$N NMB200_RANGE_UPDATE=$mode exit $?" | tee -a "$out/summary.txt"
  python - "$out/bench_ende_n${N}_range$mode.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("   value", round(d["ms_per_step"], 3), "ms", round(d["value"]), "tok/s | e2e", round(d["e2e"]["ms_per_step"], 3),
          "ms | host per step", round(d["host_enqueue_ms_per_step"], 2), "ms | waited for the exchange", d.get("exposed_comm_ms_per_step"),
          "| clocks", d.get("clocks"))
except Exception as e:
    print("    unreadable", e)
PY
done
