#!/bin/bash
set -u
out=gpurun_out/r04_trace; mkdir -p $out
timeout 200 python tools/gemm_trace.py > $out/gemm_trace.txt 2>&1; echo "trace exit $?"; cat $out/gemm_trace.txt | tail -60
