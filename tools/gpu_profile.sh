#!/bin/bash
# One gpurun call: parity suite, then the ncu launch list and one full capture of the top kernel.
# usage: tools/gpu_profile.sh <tag> [bench flags...]
set -u
tag=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest.log
[ $rc -eq 0 ] || exit 1
B="python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --plugin-seconds 0 $*"
$B > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu1_$tag.log 2>&1
echo "launch list rc=$?"
$B > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decimate -s 4 -c 2 -f -o gpurun_out/prof_$tag $B > gpurun_out/ncu2_$tag.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
