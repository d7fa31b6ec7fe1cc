"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total and mean time, share."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
d = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0][:64]; v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    d.setdefault(k, []).append(v)
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print("%-66s n=%4d total=%9.1f us mean=%8.1f  %5.1f %%" % (k, len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
print("total %.1f us in %d launches" % (tot, sum(len(v) for v in d.values())))
