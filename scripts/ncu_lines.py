"""Per-source-line instruction / stall / local-memory totals of one kernel from an .ncu-rep (development aid).
usage: ncu_lines.py rep kernel-regex [file-filter]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; data = []
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 5 and r[0].isdigit():
        def g(name):
            try: return float(r[hdr.index(name)])
            except Exception: return 0.0
        data.append((cur, int(r[0]), r[1][:90], g("Warp Stall Sampling (All Samples)"), g("Instructions Executed"), g("Thread Instructions Executed"), g("L1 Tag Requests Global"), g("L2 Theoretical Sectors Local")))
tot_s = sum(d[3] for d in data) or 1; tot_i = sum(d[4] for d in data) or 1
print("total warp instr %.3g, stall samples %d" % (tot_i, tot_s))
# regions of kernels.cu by function (line ranges from the source itself)
import re
src = open("restir_b200/csrc/kernels.cu").read().split("\n")
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:template.*\n)?(?:RS_D|__device__|__global__|static|RS_HD)[^;]*?\b(\w+)\s*\(", l)
    if m and not l.startswith("    "): marks.append((i, m.group(1)))
def region(f, ln):
    if f != "kernels.cu": return f
    name = "?"
    for i, n in marks:
        if i <= ln: name = n
        else: break
    return name
agg = {}
for f, ln, txt, s, ins, tins, l1, loc in data:
    k = region(f, ln)
    a = agg.setdefault(k, [0, 0, 0, 0]); a[0] += s; a[1] += ins; a[2] += tins; a[3] += loc
print("%-28s %8s %8s %10s %12s" % ("region", "stall%", "instr%", "lanes/instr", "local sectors"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if a[1] / tot_i > 0.003: print("%-28s %8.2f %8.2f %10.1f %12.3g" % (k, 100 * a[0] / tot_s, 100 * a[1] / tot_i, a[2] / max(a[1], 1), a[3]))
