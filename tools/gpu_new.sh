#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_imagenet.py -m gpu -q --no-header -p no:cacheprovider --maxfail=30 --timeout=240 > gpurun_out/pytest_new.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error|error" gpurun_out/pytest_new.log | tail -60
