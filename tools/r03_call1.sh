#!/bin/bash
# Round-3 first GPU call: default suite, end-to-end segment agreement, then correctness + timing of the variants that
# had never run on hardware (SYL_STREAMK, SYL_RESID_EPI, SYL_CONV0_MB, SYL_LN_WARPS).  Logs in gpurun_out/r03_1/.
set -u
out=gpurun_out/r03_1
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/smi.txt 2>&1
echo "== default GPU suite"
timeout 1200 python -m pytest tests -q -m gpu --deselect tests/test_gpu_agreement.py > $out/pytest_gpu.log 2>&1; echo "exit $?"; tail -n 8 $out/pytest_gpu.log
echo "== agreement"
timeout 900 python -m pytest tests/test_gpu_agreement.py -q -m gpu -s > $out/pytest_agreement.log 2>&1; echo "exit $?"; tail -n 5 $out/pytest_agreement.log
echo "== variant checks"
SYL_TEST_VARIANTS=1 timeout 1500 python -m pytest tests/test_gpu_variants.py -q -m gpu -s > $out/pytest_variants.log 2>&1; echo "exit $?"
grep -E "rel diff|moved boundary|deterministic|passed|failed|OK|FAILED|worst|variant:" $out/pytest_variants.log | tail -n 70
run() { name=$1; shift; echo "== bench $name"; env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu > $out/bench_$name.json 2> $out/bench_$name.err || echo "FAILED $name"; }
run default
run streamk SYL_STREAMK=1
run resid1 SYL_RESID_EPI=1
run resid2 SYL_RESID_EPI=2
run conv0mb5 SYL_CONV0_MB=5
run lnwarps4 SYL_LN_WARPS=4
run all SYL_STREAMK=1 SYL_RESID_EPI=2 SYL_CONV0_MB=5
run exact --mode exact
run default_again
python tools/bench_summary.py $out/bench_*.json
