#!/bin/bash
# compute-sanitizer passes over the GPU parity suite (memcheck: everything; racecheck: the shared-memory-heavy image / batch /
# detection / sequence / clique tests).  Usage (under gpurun): bash tools/sanitize.sh <tag>
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q > $O/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/${TAG}_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_image.py tests/test_gpu_batch.py tests/test_gpu_features.py tests/test_gpu_seq.py tests/test_gpu_klt.py tests/test_gpu_geometry.py -m gpu -x -q > $O/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/${TAG}_racecheck.log
( echo "== memcheck"; grep -E "passed|failed|ERROR SUMMARY|memcheck rc" $O/${TAG}_memcheck.log | tail -4; echo "== racecheck"; grep -E "passed|failed|RACECHECK SUMMARY|racecheck rc" $O/${TAG}_racecheck.log | tail -4 ) > $O/${TAG}_sanitizer.txt
cat $O/${TAG}_sanitizer.txt
