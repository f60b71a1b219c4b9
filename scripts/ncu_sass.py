"""SASS-level summary of one kernel of an ncu report: opcode histogram weighted by executed count, and hot segments.
usage: python scripts/ncu_sass.py report.ncu-rep kernel_regex [launch_index]"""
import collections, csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Source" in r)
print(rows[0][:2] if h else "")
hdr = rows[h]
ie, src, smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
data = [(r[src].strip(), int(r[ie]), int(r[smp])) for r in rows[h + 1:] if len(r) > ie and r[ie].isdigit()]
tot = sum(d[1] for d in data)
print("total warp-instructions", tot, "static SASS", len(data), "samples", sum(d[2] for d in data))
op = collections.Counter()
sm = collections.Counter()
for s, n, k in data:
    t = s.split()
    o = t[1] if t[0].startswith("@") else t[0]
    op[o.split(".")[0]] += n
    sm[o.split(".")[0]] += k
for k, v in op.most_common(28):
    print("%-12s %10d %5.1f%%   samples %d" % (k, v, 100 * v / tot, sm[k]))
print("---- segments (runs of equal executed count)")
i = 0
while i < len(data):
    j = i
    while j < len(data) and data[j][1] == data[i][1]:
        j += 1
    if (j - i) * data[i][1] > 0.01 * tot:
        print("%5d-%5d count %8d len %4d share %5.1f%% samples %6d first: %s" % (i, j, data[i][1], j - i, 100.0 * (j - i) * data[i][1] / tot,
                                                                     sum(d[2] for d in data[i:j]), data[i][0][:60]))
    i = j
