#!/bin/bash
set -u
out=gpurun_out/r04_tail; mkdir -p $out
echo "== tests"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_stages.py tests/test_gpu_e2e.py tests/test_gpu_trim.py -x -q -m gpu > $out/pytest.log 2>&1; echo "exit $?"; tail -n 3 $out/pytest.log
echo "== trace"
timeout 200 python tools/gemm_trace.py > $out/gemm_trace_tail.txt 2>&1; grep -E "==|stores complete" $out/gemm_trace_tail.txt
echo "== A/B tail staging (default) vs previous build (variant)"
bash tools/ab_lib.sh $out sylber_b200/variant_notail.so
