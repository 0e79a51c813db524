#!/bin/bash
# First GPU call of the next round: correctness, then timing, of the kernel variants that were written without a GPU
# (SYL_STREAMK, SYL_RESID_EPI, SYL_CONV0_MB).  Everything runs under `timeout`; the stream-K kernel traps
# instead of spinning forever if its hand-off protocol is wrong.
# Usage (under gpurun, one GPU):  bash tools/variants_ab.sh      -> logs in gpurun_out/variants/
set -u
out=gpurun_out/variants
mkdir -p $out
echo "== default GPU suite"
timeout 1500 python -m pytest tests -x -q -m gpu > $out/pytest_gpu.log 2>&1; echo "exit $?"; tail -n 5 $out/pytest_gpu.log
echo "== variant checks (GEMM-level stream-K vs fp64, then each variant against the default build)"
SYL_TEST_VARIANTS=1 timeout 2400 python -m pytest tests/test_gpu_variants.py -q -m gpu -s > $out/pytest_variants.log 2>&1; echo "exit $?"
grep -E "rel diff|moved boundary|deterministic|passed|failed|OK|FAILED|worst" $out/pytest_variants.log | tail -n 60
run() { name=$1; shift; echo "== bench $name"; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $out/bench_$name.json 2> $out/bench_$name.err || echo "FAILED $name"; }
run default
run streamk SYL_STREAMK=1
run resid1 SYL_RESID_EPI=1
run resid2 SYL_RESID_EPI=2
run conv0mb5 SYL_CONV0_MB=5
run lnwarps4 SYL_LN_WARPS=4
run all SYL_STREAMK=1 SYL_RESID_EPI=2 SYL_CONV0_MB=5
run default_again
python tools/bench_summary.py $out/bench_*.json
