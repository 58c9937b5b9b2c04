"""Summarise an `ncu --csv --metrics gpu__time_duration.sum` launch list: per-kernel time, share, launches."""
import csv, sys, re
from collections import OrderedDict
path = sys.argv[1]; skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path, errors="replace")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]; ix = {n: i for i, n in enumerate(hdr)}
launches = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    if unit in ("ns", "nsecond"): v /= 1e3
    elif unit in ("ms", "msecond"): v *= 1e3
    launches.append((int(r[ix["ID"]]), r[ix["Kernel Name"]], r[ix["Grid Size"]], v))
launches = launches[skip:]
tot = sum(l[3] for l in launches)
print("launches=%d total=%.1f us" % (len(launches), tot))
agg = OrderedDict()
for _, name, grid, v in launches:
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("mvae::", "").replace("<unnamed>::", "")
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += v
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%7.1f us %5.1f%%  x%-3d %s" % (v, 100 * v / tot, c, k))
if "-v" in sys.argv:
    for i, name, grid, v in launches:
        print("%4d %8.1f us %-14s %s" % (i, v, grid, re.sub(r"\(.*", "", name)[-60:]))
