#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.err | tail -30; cat gpurun_out/bench.json | cut -c1-1500
