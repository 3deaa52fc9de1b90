"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel table kept
under profiles/ (count, total time, share).  usage: python scripts/launch_table.py in.csv out.txt "title"
With --last-of K only the last 1/K of the launches is kept (K evaluations were profiled)."""
import csv, sys
from collections import OrderedDict
args = [a for a in sys.argv[1:] if not a.startswith("--last-of")]
k = 1
for a in sys.argv[1:]:
    if a.startswith("--last-of="):
        k = int(a.split("=")[1])
src, out, title = args[0], args[1], (args[2] if len(args) > 2 else "")
rows = [r for r in csv.reader(l for l in open(src, errors="replace") if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows[1:]:
    if len(r) <= iv or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    unit = r[iu]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    launches.append((r[ik], us))
if k > 1:
    launches = launches[len(launches) - len(launches) // k:]
agg = OrderedDict()
for name, us in launches:
    c, t = agg.get(name, (0, 0.0))
    agg[name] = (c + 1, t + us)
tot = sum(t for _, t in agg.values())
with open(out, "w") as f:
    f.write("# %s\n" % title)
    for name, (c, t) in agg.items():
        f.write("%-112s x%-4d %9.1f us %5.1f%%\n" % (name[:110], c, t, 100 * t / tot))
    f.write("total %.2f ms over %d launches\n" % (tot / 1e3, len(launches)))
print(open(out).read())
