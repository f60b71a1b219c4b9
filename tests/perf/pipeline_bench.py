"""Region batch assembly (vb_region_batch) at the Conceptual-Captions batch shape: B=256, 36 regions x 2048-d.

    python tests/perf/pipeline_bench.py [--batch 256] [--iters 50] [--cpu-batches 3] [--once]

Lives under tests/ because its CPU-baseline leg times the oracle (test infrastructure).  Prints ONE JSON line:
  kernel     : device-resident detector output (a pool of batches larger than the 126 MB L2 rotates), CUDA events around
               a replayed CUDA graph of 20 back-to-back calls; algorithmic bytes = features read once (fp32 or bf16) + image_feat written once
               (+ boxes, masks, labels) -> GB/s against MEASURED_PEAKS.json hbm_gbs;
  e2e        : the same call from PINNED HOST buffers (upload inside the timed region), samples/s;
  cpu_oracle : the numpy restatement of the reference loader's arithmetic (oracle/pipeline_oracle.py - the checker, used
               here only as the timed CPU baseline) on this box's host, one thread, samples/s.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

from volta_b200.pipeline import RegionBatchAssembler
from volta_b200.synthetic import raw_region_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--regions", type=int, default=36)
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--cpu-batches", type=int, default=3)
ap.add_argument("--kernel-only", action="store_true", help="device time only (no host-buffer leg, no CPU baseline)")
ap.add_argument("--once", action="store_true", help="one eager launch per dtype combination on a cold batch (the command ncu profiles)")
args = ap.parse_args()

B, R, F = args.batch, args.regions, 2048
dev = torch.device("cuda:0")
peak = 6529.7
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
POOL = 4
KEYS = ("raw_feat", "raw_box", "image_hw", "num_boxes", "sel")
raws = [raw_region_batch(B, R, seed=10 + i, full=False) for i in range(POOL)]
out = {"op": "vb_region_batch", "workload": "B=%d, %d regions x %d-d + global first, ragged box counts" % (B, R, F),
       "hbm_peak_gbs": peak, "pool": "%d batches rotate (%.0f MB of fp32 features, > L2)" % (POOL, POOL * B * R * F * 4 / 1e6)}


if args.once:
    for k, (in_dt, out_dt) in enumerate(((torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16), (torch.float32, torch.float32))):
        asm = RegionBatchAssembler(R, F, 5, "first", out_dtype=out_dt)
        d = {key: v.to(dev) for key, v in raws[k].items()}
        d["raw_feat"] = d["raw_feat"].to(in_dt)
        torch.cuda.synchronize()
        asm(**{key: d[key] for key in KEYS})
        torch.cuda.synchronize()
    sys.exit(0)


def timed(fn, iters):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for in_dt, out_dt, tag in ((torch.float32, torch.bfloat16, "fp32->bf16"), (torch.bfloat16, torch.bfloat16, "bf16->bf16"),
                           (torch.float32, torch.float32, "fp32->fp32")):
    asm = RegionBatchAssembler(R, F, 5, "first", out_dtype=out_dt)
    devs = [{k: v.to(dev) for k, v in r.items()} for r in raws]
    for d in devs:
        d["raw_feat"] = d["raw_feat"].to(in_dt)
    live = sum(float((r["num_boxes"].long() - (r["sel"] == 1).sum(1)).sum()) for r in raws) / POOL   # rows actually read
    esz_in, esz_out = (4 if in_dt == torch.float32 else 2), (4 if out_dt == torch.float32 else 2)
    nbytes = live * F * esz_in + B * (R + 1) * F * esz_out + B * R * (16 + 1 + 8) + B * (R + 1) * (20 + 8)
    # device time: POOL x 5 back-to-back calls captured in one CUDA graph (no host time between the launches)
    from volta_b200 import ops
    outs = asm(**{k: devs[0][k] for k in KEYS})
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(POOL * 5):
                d = devs[i % POOL]
                ops.region_batch(d["raw_feat"], d["raw_box"], d["image_hw"], d["num_boxes"], d["sel"], asm.global_pos, *outs)
    ms = timed(lambda i: graph.replay(), max(3, args.iters // 5)) / (POOL * 5)
    host_ms = timed(lambda i: asm(**{k: devs[i % POOL][k] for k in KEYS}), args.iters)
    out["kernel " + tag] = {"us": ms * 1e3, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak,
                            "samples_per_s": B / ms * 1e3, "us_per_call_through_the_python_class": host_ms * 1e3}
    if args.kernel_only:
        continue
    pinned = [{k: v.pin_memory() for k, v in r.items()} for r in raws]
    for p in pinned:
        p["raw_feat"] = p["raw_feat"].to(in_dt).pin_memory()
    ms = timed(lambda i: asm(device=dev, **{k: pinned[i % POOL][k] for k in ("raw_feat", "raw_box", "image_hw", "num_boxes", "sel")}), args.iters)
    out["e2e " + tag] = {"ms": ms, "samples_per_s": B / ms * 1e3, "h2d_bytes": B * R * F * esz_in + B * R * 17 + B * 12}

if args.kernel_only:
    print(json.dumps(out))
    sys.exit(0)

from oracle.pipeline_oracle import assemble_region_batch  # noqa: E402  (timed CPU baseline only)

torch.set_num_threads(1)
ts = []
for i in range(args.cpu_batches):
    r = raws[i % POOL]
    a = [r[k].numpy() for k in ("raw_feat", "raw_box", "image_hw", "num_boxes", "sel")]
    t0 = time.perf_counter()
    assemble_region_batch(*a)
    ts.append(time.perf_counter() - t0)
out["cpu_oracle"] = {"samples_per_s": B / min(ts), "cores": 1, "kind": "port",
                     "sample": "%d batches of %d, numpy fp32, best" % (args.cpu_batches, B)}
print(json.dumps(out))
