#!/bin/bash
# 8-GPU call: the 10 s workload at N = 8 and, on the same box, at N = 1 (for the scaling ratio).  gpurun_out/r03_multi8/
set -u
out=gpurun_out/r03_multi8
mkdir -p $out
run() { n=$1; name=$2; shift 2; echo "== bench $name ($n GPUs)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n "$@" > $out/bench_$name.json 2> $out/bench_$name.err || { echo FAILED; tail -5 $out/bench_$name.err; }; }
run 8 10s_8gpu --steps 20 --warmup 5
python bench.py --no-cpu > $out/bench_10s_1gpu.json 2> $out/bench_10s_1gpu.err
run 8 mixed_8gpu --workload mixed --steps 5 --warmup 3 --trim
python tools/bench_summary.py $out/bench_*.json
nproc; free -g | head -2
