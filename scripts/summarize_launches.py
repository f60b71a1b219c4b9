"""Aggregates an ncu launch list (`--metrics gpu__time_duration.sum --csv`) by kernel: launches, total time, share."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[h]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) < len(H) or r[H.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[H.index("Kernel Name")]).replace("void ", "")
    try:
        val = float(r[-1].replace(",", ""))
    except ValueError:
        continue
    unit = r[H.index("Metric Unit")]
    val *= {"us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1.0)
    agg[name][0] += 1
    agg[name][1] += val
tot = sum(v[1] for v in agg.values())
print("%d launches, %.3f ms summed (serialised, cold caches: shares are meaningful, absolutes are not)" % (sum(v[0] for v in agg.values()), tot / 1e6))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s n=%5d %10.3f ms %5.1f%%" % (k[:72], v[0], v[1] / 1e6, 100 * v[1] / tot))
