#!/bin/bash
# A/B of two builds of the library on ONE box: bench.py with sylber_b200/libsylber_b200.so, then with the variant file
# swapped in (the product path always loads libsylber_b200.so), then the original restored.
#   bash tools/ab_lib.sh OUTDIR VARIANT.so [bench args]
set -u
out=$1; var=$2; shift 2
mkdir -p $out
so=sylber_b200/libsylber_b200.so
cp $so /tmp/ab_default.so
run() { timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu "$@" 2> $out/err_$tag.txt > $out/bench_$tag.json || { echo "FAILED $tag"; tail -3 $out/err_$tag.txt; }; }
tag=default_1; run "$@"
cp $var $so; tag=variant_1; run "$@"
cp /tmp/ab_default.so $so; tag=default_2; run "$@"
cp $var $so; tag=variant_2; run "$@"
cp /tmp/ab_default.so $so
python tools/bench_summary.py $out/bench_default_1.json $out/bench_variant_1.json $out/bench_default_2.json $out/bench_variant_2.json
