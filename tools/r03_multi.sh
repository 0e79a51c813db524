#!/bin/bash
# Multi-GPU call (gpurun --gpus 4): sharded-vs-single bit-identity over NCCL at world 2 and 4, then the 10 s and the mixed
# workload at 4 GPUs and the 10 s workload at 2.  Logs in gpurun_out/r03_multi/.
set -u
out=gpurun_out/r03_multi
mkdir -p $out
nvidia-smi -L > $out/gpus.txt
echo "== sharded tests (NCCL, one GPU per rank)"
NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_gpu_sharded.py -q -m gpu -rs > $out/pytest_sharded.log 2>&1; echo "exit $?"; tail -n 6 $out/pytest_sharded.log
run() { n=$1; name=$2; shift 2; echo "== bench $name ($n GPUs)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n "$@" > $out/bench_$name.json 2> $out/bench_$name.err || { echo FAILED; tail -5 $out/bench_$name.err; }; }
run 4 10s_4gpu --steps 20 --warmup 5
run 4 mixed_4gpu --workload mixed --steps 5 --warmup 3
run 2 10s_2gpu --steps 20 --warmup 5
run 4 60s_4gpu --workload 60s --steps 10 --warmup 3
python tools/bench_summary.py $out/bench_*.json
