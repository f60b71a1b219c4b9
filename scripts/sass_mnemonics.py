"""Static SASS evidence for every kernel of libvolta_b200.so: per kernel (demangled), the count of the mnemonics that
tell which hardware path it uses - tcgen05 MMA (UTCHMMA), TMA loads / stores (UTMALDG / UTMASTG), TMEM reads (LDTM),
warp-level tensor cores (HMMA), NVSwitch multimem loads (LDGMC), cp.async (LDGSTS), ldmatrix (LDSM), 16-byte global accesses, reductions, shuffles.

    python scripts/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "volta_b200", "csrc", "_build")
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "UTCATOMSWS", "LDGMC", "HMMA", "LDSM", "LDGSTS", "LDG.E.128", "STG.E.128",
        "LDS.128", "STS.128", "RED", "ATOM", "SHFL", "MUFU", "SYNCS", "BAR"]


def main():
    objs = sorted(f for f in os.listdir(BUILD) if f.endswith(".o"))
    print("# cuobjdump -sass of every object in volta_b200/csrc/_build (sm_100a); columns = static instruction counts")
    print("%-92s %s" % ("kernel", " ".join("%9s" % k for k in KEYS)))
    for o in objs:
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, o)], capture_output=True, text=True).stdout
        cur, counts = None, collections.OrderedDict()
        for line in out.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1)
                counts[cur] = collections.Counter()
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and cur:
                op = m.group(1)
                counts[cur]["_total"] += 1
                for k in KEYS:
                    if op.startswith(k):
                        counts[cur][k] += 1
        if not counts:
            continue
        names = list(counts)
        dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        print("## %s" % o)
        for n, d in zip(names, dem):
            d = re.sub(r"\(.*", "", d).replace("vb::", "")
            c = counts[n]
            print("%-92s %s   total %d" % (d[:92], " ".join("%9d" % c[k] for k in KEYS), c["_total"]))


if __name__ == "__main__":
    sys.exit(main())
