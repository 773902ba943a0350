#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cli.py tests/test_gpu_model.py -m gpu -q --no-header -p no:cacheprovider -k "transformer_beam or lazy_adam" > gpurun_out/pytest_cli.log 2>&1
echo "exit $?"; grep -n "Error\|error\|Exception\|assert " gpurun_out/pytest_cli.log | head -30; tail -3 gpurun_out/pytest_cli.log
