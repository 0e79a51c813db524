#!/bin/bash
# Multi-GPU call (gpurun --gpus 4): sharded-vs-single bit-identity over NCCL at world 2 and 4 and the two-device test, then
# the 10 s workload at 4 and 2 GPUs and the mixed workload (reference padding and trimmed) at 4.  gpurun_out/r03_multi/
set -u
out=gpurun_out/r03_multi
mkdir -p $out
nvidia-smi -L > $out/gpus.txt
echo "== sharded tests (NCCL, one GPU per rank)"
NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_gpu_sharded.py -q -m gpu -rs > $out/pytest_sharded.log 2>&1; echo "exit $?"; tail -n 6 $out/pytest_sharded.log
run() { n=$1; name=$2; shift 2; echo "== bench $name ($n GPUs)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n "$@" > $out/bench_$name.json 2> $out/bench_$name.err || { echo FAILED; tail -5 $out/bench_$name.err; }; }
run 4 10s_4gpu --steps 20 --warmup 5
run 2 10s_2gpu --steps 20 --warmup 5
run 4 mixed_4gpu --workload mixed --steps 5 --warmup 3
run 4 mixed_trim_4gpu --workload mixed --steps 5 --warmup 3 --trim
python bench.py --no-cpu > $out/bench_10s_1gpu.json 2> $out/bench_10s_1gpu.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/bench_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1]); e=d["e2e"]
    print(f.split('/')[-1], "device %.3f ms %.2f M" % (d["ms_per_step"], d["value"]/1e6), "| e2e %.3f ms %.2f M" % (e["ms_per_step"], e["value"]/1e6), "| no gather %.3f" % e.get("ms_per_step_without_the_all_gather",0), "| hidden on device %.3f ms" % e.get("hidden_states_left_on_device",{}).get("ms_per_step",0))
PY
