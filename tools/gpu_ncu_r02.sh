#!/bin/bash
# ncu evidence for round 2 (numbers printed by a run under ncu are never bench values):
#   1. launch list of two training steps of the default bench command
#   2. --set full of every tc_gemm_kernel launch of one training step (dense / xent instances, fp16 and tf32)
#   3. --set full of the GRU recurrences and the Bahdanau kernels
#   4. --set full of the decoding kernels (fused step, vocabulary GEMM + combine, beam top-k)
out=gpurun_out/ncu_r02
mkdir -p $out
B="python bench.py --no-cpu-baseline --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $out/launches.csv \
    $B --steps 2 --warmup 1 > $out/list.log 2>&1
echo "launch list exit $?"; wc -l $out/launches.csv
# the report files travel back through gpurun_out/ (64 MiB): two small captures instead of one of every launch
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"tc_gemm_kernel<.*\(int\)2, \(int\)1>" -s 8 -c 8 -o $out/prof_tc_gemm -f \
    $B --steps 1 --warmup 1 > $out/tc.log 2>&1
echo "tc f16 capture exit $?"
ncu --set full --clock-control none --kernel-name-base demangled -k regex:"tc_gemm_kernel<.*\(int\)4, \(int\)1>" -s 60 -c 16 -o $out/prof_tc_tf32 -f \
    $B --steps 1 --warmup 1 > $out/tc32.log 2>&1
echo "tc tf32 capture exit $?"
ncu --set full --clock-control none -k regex:"gru_seq|bahdanau" -s 12 -c 10 -o $out/prof_rnn -f \
    $B --steps 1 --warmup 1 > $out/rnn.log 2>&1
echo "rnn capture exit $?"
ncu --set full --clock-control none -k regex:"attn_decoder_step|decode_combine|beam_local|beam_merge" -s 40 -c 8 \
    -o $out/prof_decode -f python tools/decode_probe.py --no-stepwise > $out/decode.log 2>&1
echo "decode capture exit $?"
# gpurun brings back at most 64 MiB: keep the raw metric pages as CSV, drop the reports
for rep in $out/*.ncu-rep; do
  ncu -i "$rep" --page raw --csv > "${rep%.ncu-rep}.raw.csv" 2>/dev/null
  rm -f "$rep"
done
ls -la $out/
