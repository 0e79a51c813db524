#!/bin/bash
# 8-GPU call: the 10 s workload at N = 8 (device, e2e with and without the exchange, hidden states left on the GPUs).
set -u
out=gpurun_out/r03_multi8
mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > $out/bench_10s_8gpu.json 2> $out/bench_10s_8gpu.err || { echo FAILED; tail -5 $out/bench_10s_8gpu.err; }
python - <<PY
import json
d=json.loads(open("$out/bench_10s_8gpu.json").read().strip().splitlines()[-1]); e=d["e2e"]
print("8 GPUs: device %.3f ms %.2f M" % (d["ms_per_step"], d["value"]/1e6), "| e2e %.3f ms %.2f M" % (e["ms_per_step"], e["value"]/1e6), "| no gather %.3f" % e.get("ms_per_step_without_the_all_gather",0), "| hidden on device %.3f ms %.2f M" % (e["hidden_states_left_on_device"]["ms_per_step"], e["hidden_states_left_on_device"]["value"]/1e6))
PY
