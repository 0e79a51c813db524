#!/bin/bash
# Round-2 A/B matrix on one B200: every new kernel against the one it replaced, through bench.py's stage timers.
# HISTORICAL: the SYL_GEMM_EPI / SYL_CONV0_IMPL / SYL_ATTN_IMPL switches selected kernels that were removed after this
# script produced profiles/r02_bench_ab.md (commit "Remove superseded kernels"); check that commit out to re-run it.
# Usage (under gpurun): bash tools/r02_ab.sh   -> logs in gpurun_out/r02_ab/
set -u
out=gpurun_out/r02_ab
mkdir -p $out
run() { name=$1; shift; echo "== $name"; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu > $out/bench_$name.json 2> $out/bench_$name.err || echo "FAILED $name"; tail -c 300 $out/bench_$name.err; }
run new
run epi0 SYL_GEMM_EPI=0
run conv0_ffma SYL_CONV0_IMPL=0
run attn6 SYL_ATTN_IMPL=6
for st in 0 300 1200 2400; do
  echo "== attn7 stagger $st"; SYL_ATTN_STAGGER=$st python tools/attn_bench.py > $out/attn7_stagger$st.txt 2>&1; cat $out/attn7_stagger$st.txt
done
echo "== attn7 stagger 600 (default)"; python tools/attn_bench.py | tee $out/attn7_stagger600.txt
echo "== attn6"; SYL_ATTN_IMPL=6 python tools/attn_bench.py | tee $out/attn6.txt
echo "== attn7 no exp"; SYL_ATTN_DEBUG=2 python tools/attn_bench.py | tee $out/attn7_noexp.txt
echo "== attn7 no exp no P store"; SYL_ATTN_DEBUG=6 python tools/attn_bench.py | tee $out/attn7_noexp_nop.txt
python tools/e2e_streams.py | tee $out/e2e_streams.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_ab/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    st = d["stages"]
    print(f.split("bench_")[1][:-5].ljust(12), "ms/step %.3f" % d["ms_per_step"], "e2e %.3f" % d["e2e"]["ms_per_step"],
          " ".join(f"{k[:6]}={v['ms_per_step']:.3f}" for k, v in st.items()))
PY
