#!/bin/bash
# GPU call: the whole GPU suite, then bench.py (default workload).  Usage: bash tools/r03_suite.sh TAG [extra bench args]
set -u
tag=${1:-x}; shift || true
out=gpurun_out/r03_$tag
mkdir -p $out
echo "== GPU suite"
timeout 1500 python -m pytest tests -q -m gpu > $out/pytest_gpu.log 2>&1; echo "exit $?"; tail -n 12 $out/pytest_gpu.log
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 5 "$@" > $out/bench.json 2> $out/bench.err || { echo "bench FAILED"; tail -5 $out/bench.err; }
python tools/bench_summary.py $out/bench.json
