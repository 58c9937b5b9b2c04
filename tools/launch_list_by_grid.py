"""Summarise an ncu launch list (gpu__time_duration.sum CSV) per (kernel, grid) for the LAST step in the file."""
import csv, sys, collections, re
rows = []
for r in csv.reader(open(sys.argv[1], errors="replace")):
    if len(r) > 14 and r[0].isdigit():
        rows.append(r)
# columns: ID, PID, Process, Host, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric, Unit, Value
names = [r[4] for r in rows]
# last step = after the last-but-one adam_kernel
adam = [i for i, n in enumerate(names) if n.startswith("adam_kernel")]
lo = adam[-2] + 1 if len(adam) >= 2 else 0
hi = adam[-1] + 1 if adam else len(rows)
agg = collections.OrderedDict()
tot = 0.0
for r in rows[lo:hi]:
    name = re.sub(r"\(.*", "", r[4])
    v = float(r[-1].replace(",", "")) / (1000.0 if r[-2] in ("ns", "nsecond") else 1.0)
    key = (name[:48], r[8])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1; a[1] += v; tot += v
print("launches=%d total=%.1f us" % (hi - lo, tot))
for (n, g), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%8.1f us %5.1f%% x%-3d %-48s grid %s" % (t, 100 * t / tot, c, n, g))
