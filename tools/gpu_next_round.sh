#!/bin/bash
# First GPU call of the next round: everything that was written after this round's GPU budget was spent.
#   gpurun --timeout 1500 -- 'bash tools/gpu_next_round.sh'
# 1. the step-wise model variants (Nematus GRU, conditional GRU, nematus/mlp outputs) against the oracle
# 2. the fp16-operand vocabulary projection: cast kernels, kind::f16 GEMM, fused forward/backward
# 3. the bench with and without the fp16 path (same seed, same batches): loss must agree to 1e-3
# Results land in gpurun_out/next_round/.
set -u
out=gpurun_out/next_round
mkdir -p "$out"
export PYTHONUNBUFFERED=1
NMB200_UNVERIFIED=1 timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -x > "$out/variants.log" 2>&1
echo "variants exit $?" | tee -a "$out/summary.txt"
NMB200_XENT16=1 timeout 600 python -m pytest tests/test_gpu_xent16.py -m gpu -q > "$out/xent16.log" 2>&1
echo "xent16 exit $?" | tee -a "$out/summary.txt"
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > "$out/bench_tf32.json" 2> "$out/bench_tf32.log"
echo "bench tf32 exit $?" | tee -a "$out/summary.txt"
NMB200_XENT16=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > "$out/bench_xent16.json" 2> "$out/bench_xent16.log"
echo "bench xent16 exit $?" | tee -a "$out/summary.txt"
tail -n 3 "$out/variants.log" "$out/xent16.log"
cat "$out/bench_tf32.json" "$out/bench_xent16.json"
