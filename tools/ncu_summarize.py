"""Summarise `ncu --set full` raw CSV (ncu -i fwd.ncu-rep --page raw --csv) of one warm forward into a markdown table
(stdout) and profiles/ncu_traffic.json (DRAM bytes per launch of the dominant kernel).
    python tools/ncu_summarize.py gpurun_out/r03_ev/ncu_forward_raw.csv [--write-traffic]"""
import csv, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}


def val(r, name, scale_unit=None):
    i = col[name]
    v = float(r[i]) if r[i] not in ("", "n/a") else 0.0
    u = units[i]
    if scale_unit == "bytes":
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    if scale_unit == "us":
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}[u]
    return v


def short(name):
    name = name.replace("syl::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    return re.sub(r"\(.*", "", name)


# label the gemm3 launches by their position in the forward: conv1..6, proj, then per layer qkv, out, ffn1, ffn2
gemm_labels = ["conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "feature projection"]
for l in range(64):
    gemm_labels += [f"QKV (layer {l})", f"out-proj (layer {l})", f"FFN1 (layer {l})", f"FFN2 (layer {l})"]
g = 0
lines, agg, total_us = [], {}, 0.0
gemm_bytes = []
for k, r in enumerate(data):
    name = short(r[col["Kernel Name"]])
    us = val(r, "gpu__time_duration.sum", "us")
    rd, wr = val(r, "dram__bytes_read.sum", "bytes"), val(r, "dram__bytes_write.sum", "bytes")
    label = name
    if name.startswith("gemm3_tc_kernel"):
        label = f"gemm3_tc_kernel: {gemm_labels[g]}"
        g += 1
        gemm_bytes.append((gemm_labels[g - 1], rd + wr, us))
    total_us += us
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += us
    a[2] += rd + wr
    lines.append((k, label, r[col["Grid Size"]], us, val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                  val(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"), val(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                  rd / 1e6, wr / 1e6, (rd + wr) / us / 1e6 if us else 0.0, int(val(r, "launch__registers_per_thread"))))

print(f"## Launch list: one forward = {len(data)} launches, {total_us / 1e3:.2f} ms of kernel time (cold-cache, serialised under ncu)\n")
print("| kernel | launches | total us | avg us | share | DRAM MB per launch |")
print("|---|---|---|---|---|---|")
for name, (n, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {us:.1f} | {us / n:.1f} | {100 * us / total_us:.1f} % | {by / n / 1e6:.1f} |")
print("\n## `--set full` metrics, launch by launch (first encoder layer shown; the other layers repeat it)\n")
print("| # | launch | grid | us | tensor pipe active | issue active | XU (MUFU) | DRAM read / write MB | DRAM TB/s | regs |")
print("|---|---|---|---|---|---|---|---|---|---|")
shown_layers = 0
for k, label, grid, us, tens, issue, xu, rd, wr, tbs, regs in lines:
    m = re.search(r"layer (\d+)", label)
    if m and int(m.group(1)) > 0:
        continue
    if label.startswith(("layernorm_rows_kernel<768>", "attention")) and k > 24:
        continue
    print(f"| {k} | `{label}` | {grid} | {us:.1f} | {tens:.1f} % | {issue:.1f} % | {xu:.1f} % | {rd:.0f} / {wr:.0f} | {tbs:.2f} | {regs} |")
if "--write-traffic" in sys.argv and gemm_bytes:
    out = {"gemm3_tc_kernel.all_launches": {
        "launches": len(gemm_bytes),
        "dram_bytes_per_launch_mean": sum(b for _, b, _ in gemm_bytes) / len(gemm_bytes),
        "dram_bytes_total": sum(b for _, b, _ in gemm_bytes),
        "source": f"{os.path.relpath(path, ROOT)} (ncu --set full, warm second forward, batch 32 x 10 s, default mode): "
                  "dram__bytes_read.sum + dram__bytes_write.sum summed over the kernel's launches"},
        "gemm3_tc_kernel.per_launch": {lab: {"dram_bytes": b, "us_under_ncu": us} for lab, b, us in gemm_bytes if "layer" not in lab or "layer 0" in lab}}
    json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print("\nwrote profiles/ncu_traffic.json", file=sys.stderr)
