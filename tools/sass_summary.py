"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libsylber_b200.so -> stdout (profiles/sass_summary.md).
    python tools/sass_summary.py"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "sylber_b200", "libsylber_b200.so")
txt = subprocess.check_output(["cuobjdump", "-sass", so], text=True)
pats = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA", "SYNCS", "MUFU.EX2", "DFMA"]
print("| kernel | SASS instructions | " + " | ".join(pats) + " |")
print("|---" * (len(pats) + 2) + "|")
rows = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    d = subprocess.check_output(["c++filt", name], text=True).strip()
    d = re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", "").replace("void ", "").replace("syl::", ""))
    rows.append((len(re.findall(r"/\*[0-9a-f]{4}\*/", f)), d, [len(re.findall(r"\b" + re.escape(p), f)) for p in pats]))
for n, d, c in sorted(rows, reverse=True):
    print(f"| `{d}` | {n} | " + " | ".join(str(x) for x in c) + " |")
