#!/bin/bash
# 2-GPU sanity of the final code (gpurun --gpus 2): sharded bit-identity tests over NCCL, the 10 s workload on 2 GPUs, the
# trimmed mixed workload on 1 GPU (corrected FLOP accounting).  gpurun_out/r04_multi2/
set -u
out=gpurun_out/r04_multi2
mkdir -p $out
echo "== sharded tests (NCCL, one GPU per rank)"
NCCL_DEBUG=WARN timeout 600 python -m pytest tests/test_gpu_sharded.py -q -m gpu -rs > $out/pytest_sharded.log 2>&1; echo "exit $?"; tail -n 5 $out/pytest_sharded.log
echo "== bench 10 s, 2 GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_10s_2gpu.json 2> $out/bench_10s_2gpu.err || { echo FAILED; tail -5 $out/bench_10s_2gpu.err; }
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $out/bench_reference_2gpu.json 2> $out/bench_reference_2gpu.err || { echo FAILED ref; tail -5 $out/bench_reference_2gpu.err; }
tail -c 600 $out/bench_reference_2gpu.json
echo "== bench mixed trimmed, 1 GPU"
timeout 600 python bench.py --workload mixed --steps 5 --trim --no-cpu > $out/bench_mixed_trim.json 2> $out/bench_mixed_trim.err || { echo FAILED; tail -5 $out/bench_mixed_trim.err; }
python - <<PY
import json,glob
for f in ["$out/bench_10s_2gpu.json", "$out/bench_mixed_trim.json"]:
    d=json.loads(open(f).read().strip().splitlines()[-1]); e=d["e2e"]; r=d["roofline"]
    print(f.split('/')[-1], "device %.3f ms %.2f M" % (d["ms_per_step"], d["value"]/1e6), "| e2e %.3f ms %.2f M" % (e["ms_per_step"], e["value"]/1e6), "| roofline", r["frac"], r["attn_mlp_path"]["frac"], (r.get("attn_mlp_path_in_graph") or {}).get("frac"), r["step"]["frac"])
PY
