#!/bin/bash
# GPU call of round 4: the GPU suite, the segmentation stage alone, then bench lines.  Usage: bash tools/r04_call.sh TAG [suite|quick]
set -u
tag=${1:-x}; what=${2:-suite}
out=gpurun_out/r04_$tag
mkdir -p $out
echo "== GPU tests ($what)"
if [ "$what" = quick ]; then
  timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py tests/test_gpu_trim.py -x -q -m gpu > $out/pytest_gpu.log 2>&1; echo "exit $?"
else
  timeout 1500 python -m pytest tests -q -m gpu -rs > $out/pytest_gpu.log 2>&1; echo "exit $?"
fi
tail -n 8 $out/pytest_gpu.log
echo "== segmentation stage"
timeout 300 python tools/seg_bench.py 32 10 50 2>&1 | tail -2
timeout 300 python tools/seg_bench.py 8 60 20 2>&1 | tail -2
b() { name=$1; shift; timeout 900 python bench.py "$@" > $out/bench_$name.json 2> $out/bench_$name.err || { echo "FAILED $name"; tail -3 $out/bench_$name.err; }; }
b 10s --no-cpu
b 60s --workload 60s --steps 10 --no-cpu
b mixed --workload mixed --steps 5 --no-cpu
b mixed_trim --workload mixed --steps 5 --trim --no-cpu
python tools/bench_summary.py $out/bench_10s.json $out/bench_60s.json $out/bench_mixed.json $out/bench_mixed_trim.json
