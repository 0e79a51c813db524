#!/bin/bash
# Final 1-GPU call of round 3: the whole GPU suite, then every bench configuration.  gpurun_out/r03_final/
set -u
out=gpurun_out/r03_final
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/smi.txt 2>&1
echo "== GPU suite"
timeout 1500 python -m pytest tests -q -m gpu -rs > $out/pytest_gpu.log 2>&1; echo "exit $?"; tail -n 8 $out/pytest_gpu.log
b() { name=$1; shift; echo "== bench $name"; timeout 900 python bench.py "$@" > $out/bench_$name.json 2> $out/bench_$name.err || { echo "FAILED $name"; tail -3 $out/bench_$name.err; }; }
b 10s --library-baseline
b reference --impl reference --steps 5 --warmup 1
b 60s --workload 60s --steps 10
b mixed --workload mixed --steps 5
b mixed_trim --workload mixed --steps 5 --trim --no-cpu
b exact --mode exact --no-cpu
b parity --mode parity --no-cpu
b 12layers --layers 12 --no-cpu
python tools/bench_summary.py $out/bench_10s.json $out/bench_60s.json $out/bench_mixed.json $out/bench_mixed_trim.json $out/bench_exact.json $out/bench_parity.json $out/bench_12layers.json
python -c "
import json
d=json.loads(open('$out/bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', d['value'], d['cpu_baseline'])
d=json.loads(open('$out/bench_10s.json').read().strip().splitlines()[-1]); print('agreement', d['config']['segment_agreement']['clips_with_identical_segments'], 'cpu', d['cpu_baseline']['value'], 'roofline', d['roofline']['frac'], d['roofline']['attn_mlp_path']['frac'], d['roofline']['step']['frac'], 'lib', d.get('gpu_library_baseline'))
"
