#!/bin/bash
# Final 1-GPU call of round 4: GPU suite, every bench configuration, ncu launch list of a warm forward, --set full of the
# segmentation launches.  gpurun_out/r04_final/
set -u
out=gpurun_out/r04_final
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
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_all.csv python tools/ncu_forward.py > $out/ncu_forward.log 2>&1
per=$(grep -o "launches per forward: [0-9]*" $out/ncu_forward.log | grep -o "[0-9]*$")
total=$(grep -c "gpu__time_duration.sum" $out/launches_all.csv)
echo "launches total $total, per forward $per"
echo "== --set full of the segmentation launches of the warm forward"
timeout 600 ncu --set full --import-source on --clock-control none -s $((total - 3)) -c 3 -o $out/seg python tools/ncu_forward.py > $out/ncu_seg.log 2>&1
ncu -i $out/seg.ncu-rep --page raw --csv > $out/ncu_seg_raw.csv 2>/dev/null
ls -la $out/seg.ncu-rep; wc -l $out/ncu_seg_raw.csv
echo "== memcheck: segmentation tests + smoke"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "segmentation" > $out/sanitizer_memcheck_segmentation.log 2>&1; echo "exit $?"; tail -3 $out/sanitizer_memcheck_segmentation.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > $out/sanitizer_memcheck_smoke.log 2>&1; echo "exit $?"; tail -3 $out/sanitizer_memcheck_smoke.log
