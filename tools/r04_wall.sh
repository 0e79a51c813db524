#!/bin/bash
# Which part of the epilogue slows the main loop?  gemm_trace with the diagnostic switch SYL_GEMM_EPI_SKIP:
# 0 full epilogue, 4 staged in shared memory but never stored, 1 TMEM read + math only, 2 accumulator released unread.
set -u
out=gpurun_out/r04_wall; mkdir -p $out
for v in 0 4 1 2; do
  SYL_GEMM_EPI_SKIP=$v timeout 120 python tools/gemm_trace.py > $out/trace_skip$v.txt 2>&1
  echo "== SYL_GEMM_EPI_SKIP=$v"; grep -E "^== (QKV|FFN2)|tile [23]:" $out/trace_skip$v.txt | head -6
done
