"""Groups an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name and grid: count, mean / total duration."""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, ig, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
seq = []
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = r[ik].split("(")[0]
    seq.append((int(r[iid]), name, r[ig], float(r[iv].replace(",", "")) / 1000.0))
# launches in order, so that frames / strips can be told apart
for i, name, grid, us in seq:
    print("%5d %-40s %-18s %9.1f us" % (i, name[:40], grid, us))
