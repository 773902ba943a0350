#!/bin/bash
# usage: tools/gpu_run.sh [tests] [bench] [smoke]
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests)
  timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --maxfail=30 --timeout=180 > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -40;;
smoke)
  python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log;;
bench)
  python bench.py --steps 10 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
  grep "^#" gpurun_out/bench.err | head -44; grep -v "^#" gpurun_out/bench.err | tail -5
  python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print('value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'loss',d['e2e']['last_loss']);print('roofline',d['roofline']);print('clocks',d['clocks'],'launches',d['gpu_launches'])";;
esac
done
