#!/bin/bash
# One gpurun call: GPU parity suite, smoke, then bench variants (no profiler). Outputs under gpurun_out/.
# usage: tools/gpu_check.sh "name:flags" "name:flags" ...   (default: the production kernel only)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
[ $# -eq 0 ] && set -- "default:"
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  timeout 600 python bench.py $flags > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?"
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_{n}.log").read().strip().splitlines()[-1])
    print(n, "value %.0f Mbit/s kernel_ms %.4f frac %.3f e2e %.0f clocks %s" % (d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"]))
except Exception as e:
    print(n, "no json", e)
PY
done
