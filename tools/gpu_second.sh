#!/bin/bash
mkdir -p gpurun_out
echo "=== default config gemm tests"
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q --no-header -p no:cacheprovider --timeout=120 > gpurun_out/test_gpu_gemm.log 2>&1
echo "exit $?"; tail -n 8 gpurun_out/test_gpu_gemm.log
if ! grep -q " passed" gpurun_out/test_gpu_gemm.log || grep -q "failed" gpurun_out/test_gpu_gemm.log; then
  echo "=== probing MN-major descriptor space"
  timeout 900 python tools/mn_probe.py 2>&1 | tee gpurun_out/mn_probe.log
fi
for f in test_gpu_beam test_gpu_ops test_gpu_model; do
  echo "=== $f"
  timeout 600 python -m pytest tests/$f.py -m gpu -q --no-header -p no:cacheprovider --maxfail=40 --timeout=120 > gpurun_out/$f.log 2>&1
  echo "exit $?"; tail -n 14 gpurun_out/$f.log
done
