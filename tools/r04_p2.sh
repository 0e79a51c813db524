#!/bin/bash
set -u
out=gpurun_out/r04_p2; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py tests/test_gpu_trim.py tests/test_gpu_agreement.py -x -q -m gpu > $out/pytest.log 2>&1; echo "exit $?"; tail -n 4 $out/pytest.log
timeout 200 python tools/seg_bench.py 32 10 50 2>&1 | tail -1
timeout 600 python bench.py --workload mixed --steps 5 --no-cpu > $out/bench_mixed.json 2> $out/bench_mixed.err || tail -5 $out/bench_mixed.err
python tools/bench_summary.py $out/bench_mixed.json
