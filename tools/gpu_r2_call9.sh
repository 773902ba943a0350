#!/bin/bash
# Round 2, GPU call 9: in-kernel dropout (K15), pair tests again, the diagnostic GEMM shapes, Transformer step with
# the fused dropout, ncu captures of the tcgen05 instances (CSV pages)
set -u
out=gpurun_out/call9
mkdir -p "$out"
export PYTHONUNBUFFERED=1
timeout -k 10 600 python -m pytest tests/test_gpu_dropout.py tests/test_gpu_gemm_pair.py -m gpu -q > "$out/tests_new.log" 2>&1
echo "new tests exit $?" | tee -a "$out/summary.txt"
tail -n 25 "$out/tests_new.log" | cut -c1-220
timeout -k 10 900 python -m pytest tests/test_gpu_transformer.py tests/test_gpu_reference_inis.py tests/test_gpu_variants.py -m gpu -q > "$out/tests_model.log" 2>&1
echo "model tests exit $?" | tee -a "$out/summary.txt"
tail -n 6 "$out/tests_model.log" | cut -c1-220
timeout -k 10 300 python tools/gemm_sweep.py --reps 40 --set diag > "$out/sweep_diag.txt" 2> "$out/sweep_diag.err"
echo "diag sweep exit $?" | tee -a "$out/summary.txt"
cat "$out/sweep_diag.txt"
timeout -k 10 600 python bench_workloads.py transformer --no-cpu > "$out/transformer.json" 2> "$out/transformer.log"
echo "transformer exit $?" | tee -a "$out/summary.txt"
cut -c1-400 "$out/transformer.json"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file "$out/transformer_launches.csv" \
    python bench_workloads.py transformer --no-cpu > "$out/transformer_list.log" 2>&1
echo "transformer list exit $?" | tee -a "$out/summary.txt"
timeout -k 10 1500 bash tools/gpu_ncu_r02.sh > "$out/ncu.log" 2>&1
echo "ncu exit $?" | tee -a "$out/summary.txt"
tail -n 12 "$out/ncu.log"
du -sh gpurun_out
