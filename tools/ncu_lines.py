"""Attribute ncu per-SASS-instruction samples to CUDA source lines (needs -lineinfo).
usage: ncu_lines.py <report.ncu-rep> <kernel-substring> [top]"""
import csv, re, subprocess, sys, os, tempfile
from collections import defaultdict

# usage: ncu_lines.py <report> <substring of the (mangled) kernel symbol> [top] [substring of the demangled name in the report]
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
name_filter = sys.argv[4] if len(sys.argv) > 4 else None
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "chalkydri_b200", "libchalkydri_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
out, cur, on = [], None, False
for l in dis:
    if l.startswith(".text."):
        if on:
            break
        on = kern in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        out.append((cur, m.group(2)))
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.split("\n")))
# the CSV holds one section per profiled kernel: "Kernel Name",<name> / header / instruction rows
sections, cur_name, hdr, body = [], None, None, []
for r in rows:
    if r and r[0] == "Kernel Name":
        if cur_name is not None:
            sections.append((cur_name, hdr, body))
        cur_name, hdr, body = r[1] if len(r) > 1 else "", None, []
    elif cur_name is not None and hdr is None and r:
        hdr = r
    elif cur_name is not None and r:
        body.append(r)
if cur_name is not None:
    sections.append((cur_name, hdr, body))
short = name_filter if name_filter else kern.split("_kernel")[0]
sec = [x for x in sections if short in x[0]]
if not sec:
    sys.exit("kernel not found in report: " + ", ".join(x[0][:40] for x in sections))
def _total(x):
    hh = x[1]
    ii = hh.index("Instructions Executed")
    return sum(int(r[ii] or 0) for r in x[2] if len(r) > ii and (r[ii] or "0").isdigit())
sec.sort(key=_total)
_, h, sass = sec[-1]                  # the launch that did the most work
iS, iI = h.index("# Samples"), h.index("Instructions Executed")
sass = [r for r in sass if len(r) > iI]
print("sass instrs: disasm", len(out), "ncu", len(sass))
agg = defaultdict(lambda: [0, 0])
for (cur, ins), r in zip(out, sass):
    agg[cur][0] += int(r[iS] or 0)
    agg[cur][1] += int(r[iI] or 0)
tot = sum(v[0] for v in agg.values()) or 1
toti = sum(v[1] for v in agg.values()) or 1
src = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, ln = k if k else ("?", 0)
    p = os.path.join(root, "chalkydri_b200", "csrc", f)
    if f not in src and os.path.exists(p):
        src[f] = open(p).read().split("\n")
    t = src[f][ln - 1].strip()[:100] if f in src and 0 < ln <= len(src[f]) else ""
    print(f"{100*v[0]/tot:5.1f}% samples {100*v[1]/toti:5.1f}% inst  {f}:{ln}: {t}")
