#!/bin/bash
# One GPU-box visit: parity suite, bench lines, ncu launch list and full captures of the top kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
python -m pytest tests -m gpu -x -q > $O/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_gpu_tests.log
tail -3 $O/${TAG}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 600 $O/${TAG}_bench.err
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --legs none > $O/${TAG}_bench_20steps.json 2>> $O/${TAG}_bench.err   # the driver's flags
python bench.py --steps 25 --warmup 2 --mds 1 --no-cpu-baseline --legs none > $O/${TAG}_bench_mds.json 2>> $O/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --passes-per-step 2 --no-cpu-baseline > $O/${TAG}_ncu_launch.log 2>&1
for k in ${NCU_KERNELS:-k_scan16_to_l0l1 k_klt k_clique}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $O/${TAG}_full_$k \
    python bench.py --steps 1 --warmup 1 --passes-per-step 1 --no-cpu-baseline --legs none > $O/${TAG}_ncu_$k.log 2>&1
done
python tests/perf/fmt_bench.py > $O/${TAG}_fmt_bench.json 2>> $O/${TAG}_bench.err
python tools/klt_stress.py > $O/${TAG}_klt_stress.json 2>> $O/${TAG}_bench.err
python tools/ingest_bench.py > $O/${TAG}_ingest.json 2>> $O/${TAG}_bench.err
python tests/perf/sequential_bench.py > $O/${TAG}_sequential.json 2>> $O/${TAG}_bench.err
cat $O/${TAG}_bench.json
