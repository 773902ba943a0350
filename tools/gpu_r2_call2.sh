#!/bin/bash
# Round 2, GPU call 2: the fused decoder step + decoding engine, the un-gated variants, the fp16 vocabulary path.
set -u
out=gpurun_out/call2
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q -x -k "step_kernel" > "$out/decode_op.log" 2>&1
echo "decode op exit $?" | tee -a "$out/summary.txt"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_decode.py -m gpu -q -x -k "step_kernel_against_fp64 and dims1" > "$out/decode_memcheck.log" 2>&1
echo "decode memcheck exit $?" | tee -a "$out/summary.txt"
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q -k "not step_kernel" > "$out/decode_engine.log" 2>&1
echo "decode engine exit $?" | tee -a "$out/summary.txt"
NMB200_UNVERIFIED=1 timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -q > "$out/variants.log" 2>&1
echo "variants exit $?" | tee -a "$out/summary.txt"
NMB200_XENT16=1 timeout 600 python -m pytest tests/test_gpu_xent16.py -m gpu -q > "$out/xent16.log" 2>&1
echo "xent16 exit $?" | tee -a "$out/summary.txt"
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_golden.py tests/test_gpu_beam.py -m gpu -q > "$out/model.log" 2>&1
echo "model/golden/beam exit $?" | tee -a "$out/summary.txt"
timeout 600 python tools/decode_probe.py > "$out/decode_probe.json" 2> "$out/decode_probe.log"
echo "decode probe exit $?" | tee -a "$out/summary.txt"
for f in decode_op decode_memcheck decode_engine variants xent16 model; do echo "== $f"; tail -n 4 "$out/$f.log"; done
cat "$out/decode_probe.json"
