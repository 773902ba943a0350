#!/bin/bash
# 2-GPU call: the data-parallel equivalence test only (eager with the early exchange; whole step incl. NCCL in one graph)
set -u
out=gpurun_out/dp2
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -s > "$out/dp_tests.log" 2>&1
echo "dp tests exit $?" | tee -a "$out/summary.txt"
tail -n 12 "$out/dp_tests.log" | cut -c1-250
