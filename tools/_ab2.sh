for v in hi lo hi lo; do RADARFE_TAIL_PRIO=$v timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --legs none 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), round(d['e2e']['value']), {k:round(v['ms'],3) for k,v in d['stages'].items() if k in ('scan_to_l0l1','pyr_down','klt','reject')})"; done
timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --legs strong,mds > gpurun_out/t45.json 2> gpurun_out/t45.err; tail -2 gpurun_out/t45.err
python -c "
import json; d=json.load(open('gpurun_out/t45.json')); print(round(d['value']))
for k in ('strong','mds'): print(k, {a:b for a,b in d[k].items() if a!='workload'})"
