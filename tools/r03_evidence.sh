#!/bin/bash
# Evidence call (one GPU): ncu launch list + --set full metrics of every launch of one warm forward (default mode),
# compute-sanitizer memcheck and racecheck.  Outputs in gpurun_out/r03_ev/.
set -u
out=gpurun_out/r03_ev
mkdir -p $out
echo "== launch list"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_all.csv python tools/ncu_forward.py > $out/ncu_forward.log 2>&1
per=$(grep -o "launches per forward: [0-9]*" $out/ncu_forward.log | grep -o "[0-9]*$")
total=$(grep -c "gpu__time_duration.sum" $out/launches_all.csv)
skip=$((total - per))
echo "launches total $total, per forward $per, skip $skip"
echo "== --set full of the warm forward"
timeout 1200 ncu --set full --clock-control none -s $skip -c $per -o $out/fwd python tools/ncu_forward.py > $out/ncu_full.log 2>&1
ncu -i $out/fwd.ncu-rep --page raw --csv > $out/ncu_forward_raw.csv 2>/dev/null
ls -la $out/fwd.ncu-rep; wc -l $out/ncu_forward_raw.csv
rm -f $out/fwd.ncu-rep            # 118 MB: gpurun copies at most 64 MiB back; the raw CSV is what the summary reads
echo "== memcheck smoke"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > $out/sanitizer_memcheck_smoke.log 2>&1; echo "exit $?"; tail -3 $out/sanitizer_memcheck_smoke.log
echo "== racecheck smoke"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > $out/sanitizer_racecheck_smoke.log 2>&1; echo "exit $?"; tail -3 $out/sanitizer_racecheck_smoke.log
grep -E "hazard detected|at .*\+0x|by .*kernel|Race reported" $out/sanitizer_racecheck_smoke.log | sed -e 's/=========//' | sort | uniq -c | sort -rn | head -40 > $out/sanitizer_racecheck_summary.txt; head -30 $out/sanitizer_racecheck_summary.txt
echo "== memcheck kernel + trim tests"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trim.py tests/test_gpu_stages.py -q -m gpu -x > $out/sanitizer_memcheck_tests.log 2>&1; echo "exit $?"; tail -3 $out/sanitizer_memcheck_tests.log
