#!/bin/bash
mkdir -p gpurun_out
for w in "$@"; do
  timeout 400 python bench_workloads.py $w > gpurun_out/wl_$w.json 2> gpurun_out/wl_$w.err
  echo "$w exit $?"; tail -3 gpurun_out/wl_$w.err | cut -c1-300; cut -c1-1200 gpurun_out/wl_$w.json
done
