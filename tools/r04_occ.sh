#!/bin/bash
set -u
out=gpurun_out/r04_occ; mkdir -p $out
so=sylber_b200/libsylber_b200.so
cp $so /tmp/default.so
for v in default occ5 occ6 occ7; do
  [ $v = default ] && cp /tmp/default.so $so || cp sylber_b200/variant_$v.so $so
  echo "-- $v"
  timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k segmentation 2>&1 | tail -1
  timeout 200 python tools/seg_bench.py 32 10 50 2>&1 | tail -1
  timeout 200 python tools/seg_bench.py 8 60 20 2>&1 | tail -1
done
cp /tmp/default.so $so
echo "== headline bench (with the in-graph path measurement)"
timeout 600 python bench.py --library-baseline > $out/bench_10s.json 2> $out/bench_10s.err || tail -5 $out/bench_10s.err
python tools/bench_summary.py $out/bench_10s.json
python -c "
import json
d=json.loads(open('$out/bench_10s.json').read().strip().splitlines()[-1]); r=d['roofline']; print('roofline', r['frac'], 'path eager', r['attn_mlp_path']['frac'], 'path in graph', r['attn_mlp_path_in_graph'], 'step', r['step']['frac'])
"
