#!/bin/bash
# ncu evidence for the round: (1) launch list of two bench steps, (2) full capture of the
# dominant kernels.  Numbers printed by a run under ncu are never bench values.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench_list.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches.csv
ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 8 -c 6 -o gpurun_out/prof_tc_gemm -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_tc.log 2>&1
echo "tc capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:gru_seq -c 2 -o gpurun_out/prof_gru -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_gru.log 2>&1
echo "gru capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:bahdanau -c 4 -o gpurun_out/prof_att -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_att.log 2>&1
echo "att capture exit $?"
ls -la gpurun_out/*.ncu-rep
