"""Print one line per bench JSON file: python tools/bench_summary.py gpurun_out/r02_b/bench_*.json"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e)
        continue
    st = d.get("stages", {})
    print(f.split("/")[-1].ljust(24), "ms/step %.3f" % d["ms_per_step"], "value %.0f" % d["value"],
          "| e2e %.3f ms %.0f |" % (d["e2e"].get("ms_per_step", 0), d["e2e"]["value"]),
          " ".join(f"{k[:6]}={v['ms_per_step']:.3f}" for k, v in st.items()))
