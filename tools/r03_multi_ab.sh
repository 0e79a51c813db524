#!/bin/bash
# usage: bash tools/r03_multi_ab.sh N  -> bench at N GPUs with and without host-core affinity; prints the e2e split
set -u
n=$1
out=gpurun_out/r03_multi_ab$n
mkdir -p $out
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n "$@" > $out/bench_$name.json 2> $out/bench_$name.err || { echo FAILED $name; tail -5 $out/bench_$name.err; }; }
run affinity --steps 20 --warmup 5
run noaffinity --steps 20 --warmup 5 --no-affinity
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/bench_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], "device %.3f ms" % d["ms_per_step"], "| e2e %.3f ms" % d["e2e"]["ms_per_step"], "| e2e without all-gather %.3f ms" % d["e2e"].get("ms_per_step_without_the_all_gather", 0), "|", d["config"].get("host_affinity"))
PY
