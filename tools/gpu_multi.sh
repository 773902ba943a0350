#!/bin/bash
# usage: tools/gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"; grep -i "error\|warn\|fail" gpurun_out/bench_n$N.err | head -5; python -c "
import json;d=json.load(open('gpurun_out/bench_n$N.json'));print('N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'loss',d['e2e']['last_loss'],'host',d['host_enqueue_ms_per_step'])"
