#!/bin/bash
# Multi-GPU visit: all-GPU H2D probe + the bench under torchrun.  Usage (under gpurun --gpus N): bash tools/multi_gpu_round.sh <tag> <N>
TAG=${1:-rXX}; N=${2:-2}
O=gpurun_out; mkdir -p $O tools/probe
[ -x tools/probe/h2d_probe_multi.bin ] || nvcc -O3 -gencode arch=compute_100a,code=sm_100a -Xcompiler -pthread tools/h2d_probe_multi.cu -o tools/probe/h2d_probe_multi.bin
for n in $N $((N/2)) 1; do [ $n -ge 1 ] && tools/probe/h2d_probe_multi.bin $n; done > $O/${TAG}_h2d_probe_n$N.txt 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 --legs strong,chained --no-cpu-baseline > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err
tail -2 $O/${TAG}_bench_n$N.err
python -c "
import json; d=json.load(open('$O/${TAG}_bench_n$N.json')); print('N=$N value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['h2d_only'])
for k in ('strong','chained'):
    v=d.get(k,{}); print(k, round(v.get('value',0)), v.get('unit'), 'e2e', round(v.get('e2e',{}).get('value',0)))"
