"""DRAM traffic of the tcgen05 GEMM launches of ONE training step, from an ncu launch list with
gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch (feeds bench.py's roofline.traffic).
usage: python scripts/gemm_traffic.py launches.csv [step_index] > profiles/r02_gemm_traffic.json"""
import collections, csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[h]
per = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) < len(H):
        continue
    i = r[0]
    name = re.sub(r"\(.*", "", r[H.index("Kernel Name")]).replace("void ", "")
    try:
        val = float(r[-1].replace(",", ""))
    except ValueError:
        continue
    unit = r[H.index("Metric Unit")]
    metric = r[H.index("Metric Name")]
    if metric == "gpu__time_duration.sum":
        val *= {"us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1.0)
    else:
        val *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)
    per.setdefault(i, {"name": name})[metric] = val
seq = list(per.values())
marks = [i for i, d in enumerate(seq) if d["name"].startswith("emb::gather_fwd_kernel")]
lo, hi = marks[which], (marks[which + 1] if which + 1 < len(marks) else len(seq))
step = seq[lo:hi]
gemm = [d for d in step if d["name"].startswith("tc::")]
tot_t = sum(d.get("gpu__time_duration.sum", 0.0) for d in step)
g_t = sum(d.get("gpu__time_duration.sum", 0.0) for d in gemm)
g_b = sum(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in gemm)
print(json.dumps({"kernel": "vb::tc::gemm_tc_kernel", "launches_per_step": len(gemm), "launches_in_step": len(step),
                  "dram_bytes_per_step": g_b, "dram_bytes_per_launch_avg": g_b / max(1, len(gemm)),
                  "gemm_share_of_step_time": g_t / tot_t if tot_t else None, "step_ms_summed_under_ncu": tot_t / 1e6,
                  "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                            "(eager steps of bench.py, ctrl_vilbert_base B=256 T=36 V=37, packed rows); one step between two "
                            "text-embedding gathers"}, indent=1))
