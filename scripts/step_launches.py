"""One training step out of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): the launches between two
consecutive occurrences of a marker kernel (default: the text-embedding gather, first kernel of every forward), aggregated
by kernel and by family.   usage: python scripts/step_launches.py launches.csv [step_index] [marker]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 2
marker = sys.argv[3] if len(sys.argv) > 3 else "emb::gather_fwd_kernel"
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[h]
seq = []
for r in rows[h + 1:]:
    if len(r) < len(H) or r[H.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[H.index("Kernel Name")]).replace("void ", "")
    try:
        val = float(r[-1].replace(",", ""))
    except ValueError:
        continue
    val *= {"us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(r[H.index("Metric Unit")], 1.0)
    seq.append((name, val))
marks = [i for i, (n, _) in enumerate(seq) if n.startswith(marker)]
print("%d launches in the file, marker '%s' at %s" % (len(seq), marker, marks))
lo, hi = marks[which], (marks[which + 1] if which + 1 < len(marks) else len(seq))
step = seq[lo:hi]
agg = collections.defaultdict(lambda: [0, 0.0])
fam = collections.defaultdict(float)
for n, v in step:
    agg[n][0] += 1; agg[n][1] += v
    f = ("GEMM (tcgen05)" if n.startswith("tc::") else "attention" if n.startswith("attn::mma") else "LayerNorm" if n.startswith("ln::")
         else "bias-gradient column sums" if "colsum" in n else "fill / zero rows" if ("fill" in n.lower() or "zero_rows" in n) else "other")
    fam[f] += v
tot = sum(v for _, v in step)
print("step %d: %d launches, %.3f ms summed (ncu serialises and flushes caches: shares are meaningful, absolutes are not)" % (which, len(step), tot / 1e6))
for k, v in sorted(fam.items(), key=lambda kv: -kv[1]):
    print("  %-28s %8.3f ms %5.1f%%" % (k, v / 1e6, 100 * v / tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-72s n=%4d %9.3f ms %5.1f%%  avg %6.1f us" % (k[:72], v[0], v[1] / 1e6, 100 * v[1] / tot, v[1] / v[0] / 1e3))
