#!/bin/bash
set -u
out=gpurun_out/r04_seg_ab; mkdir -p $out
so=sylber_b200/libsylber_b200.so
echo "== tests (default build: chunk 32, pool grid 128)"
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -x -q -m gpu > $out/pytest.log 2>&1; echo "exit $?"; tail -n 3 $out/pytest.log
cp $so /tmp/default.so
for v in default chunk16 chunk8; do
  [ $v = default ] && cp /tmp/default.so $so || cp sylber_b200/variant_$v.so $so
  echo "-- $v"
  timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k segmentation 2>&1 | tail -1
  timeout 200 python tools/seg_bench.py 32 10 50 2>&1 | tail -1
  timeout 200 python tools/seg_bench.py 8 60 20 2>&1 | tail -1
done
cp /tmp/default.so $so
