#!/bin/bash
# Round 2, GPU call 6: phase stamps of the decoder step; variants with fused gate kernels; reference INIs
set -u
out=gpurun_out/call6
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 600 python tools/decstep_phases.py > "$out/phases.txt" 2> "$out/phases.log"
echo "phases exit $?" | tee -a "$out/summary.txt"
timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_reference_inis.py tests/test_gpu_decode.py -m gpu -q > "$out/tests.log" 2>&1
echo "tests exit $?" | tee -a "$out/summary.txt"
cat "$out/phases.txt"; tail -n 5 "$out/tests.log"
