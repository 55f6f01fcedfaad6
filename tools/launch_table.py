"""Per-kernel totals of the LAST n launches in an ncu `--metrics gpu__time_duration.sum --csv` launch list."""
import csv, collections, sys
path, n = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
names = [r[ki] for r in rows[1:]]; vals = [float(r[vi].replace(',', '')) for r in rows[1:]]
agg = collections.OrderedDict()
for k, v in zip(names[-n:], vals[-n:]):
    k = k.split('(')[0].split('::')[-1]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("last %d launches: %.1f us (cold-cache, serialised ncu times)" % (n, tot / 1e3))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-24s n=%3d total %9.1f us  avg %7.2f us  %5.1f%%" % (k, c, t / 1e3, t / c / 1e3, 100 * t / tot))
