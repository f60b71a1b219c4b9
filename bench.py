#!/usr/bin/env python
"""Headline benchmark: ctrl_vilbert_base pre-training (MLM + MRM + ITM) forward + backward, samples/s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on the host CPU

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one synthetic batch: forward of
BertForVLPreTraining in train mode (dropout on), the three losses, backward of their sum, and - with N > 1 -
the bucketed gradient all-reduce overlapped with the backward.  The optimizer is not part of the path
(SURVEY.md 8f).  Shapes follow BASELINE.json: 36 text tokens, 36 regions + global = 37 x 2048-d features,
5-d boxes, 1601-way region targets; per-GPU batch 256 (weak scaling).

  value     samples/s with the batch already resident in HBM when the timed region starts
  e2e       the same metric through the public API with HOST (pinned) buffers: volta_b200.pipeline.PrefetchingRegionLoader
            uploads every step's RAW detector output + text tensors on a copy stream (while the previous step computes),
            assembles image_feat / image_loc / image_mask / image_label on the device (vb_region_batch), and the three
            losses are read back every step; their copy is queued right after the forward, so the host blocks on the
            forward - not on the backward - before it queues the next step (all of it inside the timed region)
  roofline  dominant kernel = the tcgen05 GEMM: executed FLOPs (2*M*N*K per launch, device-side row counts applied) / CUDA-
            event time of every GEMM launch of an instrumented step, against the measured bf16 peak that matches the
            clocks seen during the timed region (burst when un-throttled, sustained when power-capped)
  cpu_baseline / --impl reference: the oracle port of the reference (oracle/volta_oracle.py, fp32, train mode)
            timed on this box's host cores on a bounded sample (batch 8 per step)

Other BASELINE.json configs (recorded under profiles/, not the driver's default line):
    --model ctrl_visualbert_base | ctrl_lxmert | ctrl_uniter_base | ctrl_vl-bert_base     (pre-training, same contract)
    --batch 32 --seq-len 38                                       (the reference recipe: global batch 256 over 8 GPUs)
    --workload retrieval [--model vilbert_base --seq-len 30 --regions 101 --pairs 500]    (config 5: forward-only scoring)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_SAMPLE = {  # algorithmic fwd+bwd GFLOP per sample of the REFERENCE algorithm
    ("ctrl_vilbert_base", 36, 37): 52.578, ("ctrl_vilbert_base", 38, 37): 54.420, ("ctrl_visualbert_base", 36, 37): 43.627,
    ("ctrl_lxmert", 36, 37): 48.537, ("ctrl_uniter_base", 36, 37): 43.628, ("ctrl_vl-bert_base", 36, 37): 44.093,
    ("vilbert_base", 30, 101): 77.646,
}
FLOPS_PER_PAIR = {("vilbert_base", 30, 101): 24.039, ("vilbert_base", 36, 37): 15.406}   # BertModel forward only


def metric_name(args):
    if args.workload == "retrieval":
        return "%s retrieval scoring pairs/s" % args.model
    return "ctrl_vilbert pretrain samples/s" if args.model == "ctrl_vilbert_base" else "%s pretrain samples/s" % args.model


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="volta_b200", choices=["volta_b200", "reference"])
    ap.add_argument("--workload", default="pretrain", choices=["pretrain", "retrieval"])
    ap.add_argument("--model", default=None)
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--seq-len", type=int, default=None)
    ap.add_argument("--regions", type=int, default=None, help="visual tokens including the global feature")
    ap.add_argument("--pairs", type=int, default=500, help="retrieval: candidate images per caption = pairs per step and GPU")
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--raw-input", default="bf16", choices=["fp32", "bf16"],
                    help="e2e leg: dtype in which the region features travel host->device (the GEMM operand is bf16 either way)")
    ap.add_argument("--no-optimizer-leg", action="store_true")
    ap.add_argument("--no-full-length-leg", action="store_true")
    ap.add_argument("--full-length", action="store_true", help="headline batch without padding (every len = T, every region valid)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    a = ap.parse_args()
    if a.model is None:
        a.model = "vilbert_base" if a.workload == "retrieval" else "ctrl_vilbert_base"
    if a.seq_len is None:
        a.seq_len = 30 if a.workload == "retrieval" else 36
    if a.regions is None:
        a.regions = 101 if a.workload == "retrieval" else 37
    return a


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(sustained=d.get("bf16_tflops_sustained", 1354.0), burst=d.get("bf16_tflops", 1651.7),
                    hbm=d.get("hbm_gbs", 6529.7), source="measured (MEASURED_PEAKS.json)")
    return dict(sustained=1400.0, burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def pick_peak(pk, clocks):
    """The denominator that matches how the kernel was timed: the burst figure when the SM clock stayed at its maximum and
    nothing capped it during the timed region, the sustained one when the sampler saw a power cap / lower clocks."""
    sm, mx = clocks.get("sm_mhz"), clocks.get("sm_max_mhz")
    if sm and mx and sm >= 0.95 * mx and "sw_power_cap" not in clocks.get("reasons", []):
        return pk["burst"], pk["source"] + ", burst: SM clock %.0f of %.0f MHz, no cap" % (sm, mx)
    return pk["sustained"], pk["source"] + ", sustained: clocks below maximum or power-capped during the timed region"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, str(gpu_index)

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.idx, "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def end(self):
        self.t1 = time.perf_counter()

    def mark(self):
        """Start of the timed region: nvidia-smi is started before the warm-up (it needs a few hundred ms to produce its
        first line - longer than a short timed region), only samples that arrive after this mark are summarised."""
        self.t0 = time.perf_counter()

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        t0 = getattr(self, "t0", 0.0)
        t1 = getattr(self, "t1", float("inf")) + 0.05
        rows = [r for ts, r in self.rows if t0 <= ts <= t1]
        if not rows and self.rows:          # region shorter than the sampling period: the sample taken closest to it
            rows = [min(self.rows, key=lambda tr: abs(tr[0] - 0.5 * (t0 + min(t1, t0 + 60.0))))[1]]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(args, world, batch=None):
    B, T, V = (args.batch if batch is None else batch), args.seq_len, args.regions
    if args.workload == "retrieval":
        return {"workload": "%s Flickr30k-shape 1:N retrieval scoring, forward only (BertForVLTasks VL-logit head, eval mode): one "
                            "caption of T=%d tokens against %d candidate images per step and GPU, V=%d regions incl. global x "
                            "2048-d + 5-d boxes; work items (caption x image chunk) sharded over the ranks, no data-path "
                            "collective" % (args.model, T, args.pairs, V),
                "pairs_per_step_per_gpu": args.pairs, "parallelism": "replicas x%d" % world,
                "l2": "working set (414 MB of fp32 features per step) exceeds the 126 MB L2; no flush needed"}
    return {"workload": "%s pre-training fwd+bwd (MLM+MRM+ITM, train mode, dropout on), per-GPU batch %d, "
                        "T=%d text tokens, V=%d regions incl. global x 2048-d + 5-d boxes%s" %
                        (args.model, B, T, V, ", no padding (--full-length)" if args.full_length else ""),
            "per_gpu_batch": B, "global_batch": B * world, "parallelism": "dp%d" % world,
            "l2": "working set (activations ~25 MB/sample, bf16 weights 484 MB) exceeds the 126 MB L2; no flush needed",
            "launch": "CUDA graphs (forward graph + backward graph segments), replayed per step",
            "mlm_head": "vocabulary decoder + cross-entropy run on the labelled (masked) rows only, count kept on the device; "
                        "the reference computes all T positions and ignores ~90 % of them (ignore_index=-1) - same loss and "
                        "gradients.  roofline.whole_step_frac_algorithmic uses the reference algorithm's FLOPs, "
                        "roofline.whole_step_frac_executed the FLOPs really issued (padding rows and unlabelled MLM rows are "
                        "skipped; VOLTA_B200_MLM_COMPACT=0 / VOLTA_B200_PACK_ROWS=0 run everything)"}


def make_cfg(args, dropout=True):
    from volta_b200.config import named_config
    over = {} if dropout else dict(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, v_hidden_dropout_prob=0.0,
                                   v_attention_probs_dropout_prob=0.0)
    return named_config(args.model, **over)


def batch_args(d):
    return (d["input_ids"], d["image_feat"], d["image_loc"], d["segment_ids"], d["input_mask"], d["image_mask"],
            d["lm_label_ids"], d["image_label"], d["image_cls"], None, None, None, None, None, d["is_match"])


RETRIEVAL_TASK = "TASK8"


def retrieval_task_cfg():
    """config_tasks/vilbert_test_tasks.yml:36-50 (RetrievalFlickr30k): a VL-logit head over the fused pooled output."""
    return {RETRIEVAL_TASK: {"type": "VL-logit", "num_labels": 1}}


def retrieval_item(args, seed, step=0):
    """One work item as RetrievalDatasetVal.__getitem__ yields it (volta/datasets/retrieval_dataset.py:377-414): one caption,
    `pairs` candidate images."""
    import torch
    from volta_b200.synthetic import pretraining_batch
    b = pretraining_batch(args.pairs, args.seq_len, args.regions, seed=seed)
    cap = pretraining_batch(1, args.seq_len, args.regions, seed=seed + 1)
    target = torch.zeros(1, args.pairs)
    target[0, step % args.pairs] = 1
    return (b["image_feat"].unsqueeze(0), b["image_loc"].unsqueeze(0), b["image_mask"].unsqueeze(0), cap["input_ids"],
            cap["input_mask"], cap["segment_ids"], target, torch.tensor([step]), torch.tensor([0]))


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference(args, steps, warmup):
    """The reference algorithm (oracle port, fp32) on the host cores, on a bounded sample of the workload."""
    import contextlib
    import torch
    from oracle import volta_oracle as O
    import volta_b200
    from volta_b200.synthetic import pretraining_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = make_cfg(args)
    torch.manual_seed(0)
    B = args.cpu_batch
    if args.workload == "retrieval":
        with contextlib.redirect_stdout(sys.stderr):
            model = volta_b200.BertForVLTasks(cfg, retrieval_task_cfg(), [RETRIEVAL_TASK])
        P = O.params_from_state_dict(model.state_dict())
        b = pretraining_batch(B, args.seq_len, args.regions, seed=1234)

        def one():
            with torch.no_grad():
                O.retrieval_scores(P, cfg, b, clf="clfs_dict.%s" % RETRIEVAL_TASK)
        what = "%s fp32 eval-mode forward of %d pairs, T=%d V=%d" % (args.model, B, args.seq_len, args.regions)
        unit = "pairs/s"
    else:
        with contextlib.redirect_stdout(sys.stderr):
            model = volta_b200.BertForVLPreTraining(cfg)  # parameter container only: supplies reference-initialised weights
        P = O.params_from_state_dict(model.state_dict(), requires_grad=True)
        b = pretraining_batch(B, args.seq_len, args.regions, seed=1234)

        def one():
            for v in {id(v): v for v in P.values()}.values():
                v.grad = None
            O.fwd_bwd(P, cfg, b, train=True)
        what = "%s fp32 train-mode fwd+bwd, batch %d, T=%d V=%d" % (args.model, B, args.seq_len, args.regions)
        unit = "samples/s"
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return dict(value=B / med, unit=unit, cores=torch.get_num_threads(), kind="port",
                sample="%s, %d warm-up + %d timed steps, median %.3f s/step" % (what, warmup, steps, med)), med


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    cb, med = cpu_reference(args, steps, warmup)
    cfg = {k: v for k, v in workload_config(args, 1, batch=args.cpu_batch).items() if k in ("workload", "per_gpu_batch", "global_batch")}
    cfg["workload"] += " - BOUNDED SAMPLE timed on the host CPU: batch %d per step (the GPU arm runs %s per GPU and step)" % (
        args.cpu_batch, "%d pairs" % args.pairs if args.workload == "retrieval" else "batch %d" % args.batch)
    cfg["parallelism"] = "host cpu, %d threads" % cb["cores"]
    line = {"impl": "reference", "metric": metric_name(args), "value": cb["value"], "unit": cb["unit"], "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": cb["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------ GPU arm: shared pieces
class Harness:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: volta_b200 has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            if os.environ.get("VOLTA_B200_NCCL_CTAS"):  # experiments with VOLTA_B200_DDP_RESERVE_SMS (parallel.py)
                os.environ.setdefault("NCCL_MAX_CTAS", os.environ["VOLTA_B200_NCCL_CTAS"])
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's banner must not share stdout with the JSON line
            dist.init_process_group(backend="nccl", device_id=self.dev)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        from volta_b200 import engine
        torch = self.torch
        with ClockSampler(self.local) as cs:
            for _ in range(warmup):
                fn()
            self.barrier()
            n0 = engine.launch_count
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            cs.mark()
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            self.barrier()
            cs.end()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps, engine.launch_count - n0, cs.summary()

    def gemm_roofline(self, run_once, ms_step, clocks, units_per_step, algo_gflop_per_unit, unit_name):
        """One instrumented EAGER pass (the timed legs may replay CUDA graphs, whose kernels cannot be bracketed one by
        one): every GEMM launch is re-issued 5x back to back between two CUDA events on the launching stream
        (vb_gemm_bench) - device time per launch without host gaps, operands as L2-warm as in the real step."""
        from volta_b200 import ops as vb_ops
        torch = self.torch
        vb_ops.GEMM_BENCH = []
        run_once()
        torch.cuda.synchronize()
        recs, vb_ops.GEMM_BENCH = vb_ops.GEMM_BENCH, None
        tc = [(f, ms) for f, ms, is_tc, _ in recs if is_tc]
        if os.environ.get("VOLTA_B200_BENCH_VERBOSE") and self.rank == 0:
            agg = {}
            for f, ms, is_tc, shape in recs:
                t = agg.setdefault((shape, is_tc), [0, 0.0, 0.0])
                t[0] += 1; t[1] += ms; t[2] += f
            for (shape, is_tc), (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                print("gemm %-44s tc=%d n=%3d  %8.3f ms  %7.1f TFLOP/s" % (shape, is_tc, n, ms, fl / ms / 1e9 if ms > 0 else 0), file=sys.stderr)
        pk = peaks()
        peak, peak_source = pick_peak(pk, clocks)
        flops_tc, ms_tc = sum(f for f, _ in tc), sum(t for _, t in tc)
        achieved = flops_tc / (ms_tc * 1e-3) / 1e12 if ms_tc > 0 else 0.0
        executed = sum(f for f, _, _, _ in recs)
        per_s = units_per_step / (ms_step * 1e-3)        # per GPU
        out = {"bound": "tensor", "kernel": "vb::tc::gemm_tc_kernel (tcgen05/TMEM/TMA)", "achieved": achieved, "peak": peak,
               "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_source,
               "what": "executed FLOPs of all tensor-core GEMM launches of one step / their summed launch durations "
                       "(CUDA events on the launching stream, each launch repeated 5x back to back)",
               "launches_per_step": len(tc), "gemm_ms_per_step": ms_tc, "gemm_share_of_step": ms_tc / ms_step,
               "whole_step_frac_executed": executed / (ms_step * 1e-3) / 1e12 / pk["burst"],
               "whole_step_frac_algorithmic": None if algo_gflop_per_unit is None else algo_gflop_per_unit * 1e9 * per_s / (pk["burst"] * 1e12),
               "whole_step_peak": "burst (%.1f TFLOP/s)" % pk["burst"],
               "algorithmic_gflop_per_%s" % unit_name: algo_gflop_per_unit,
               "executed_gemm_gflop_per_%s" % unit_name: executed / units_per_step / 1e9}
        return out


# ------------------------------------------------------------------------------------------ GPU arm: pre-training
def main_pretrain(args):
    import contextlib
    H = Harness(args)
    torch, dist, dev, world, rank = H.torch, H.dist, H.dev, H.world, H.rank
    import volta_b200
    from volta_b200.pipeline import PrefetchingRegionLoader, RegionBatchAssembler
    from volta_b200.synthetic import pretraining_batch, raw_region_batch

    cfg = make_cfg(args)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(sys.stderr):  # the constructor prints the visual targets, as the reference does
        model = volta_b200.BertForVLPreTraining(cfg).set_precision(args.precision)
    model.train().to(dev)
    runner = model
    if world > 1:
        from volta_b200.parallel import DistributedDataParallel
        runner = DistributedDataParallel(model)

    B, T, V = args.batch, args.seq_len, args.regions
    host = pretraining_batch(B, T, V, seed=1234 + rank, full_length=args.full_length)
    resident = {k: v.to(dev) for k, v in host.items()}

    def step(d):
        model.zero_grad(set_to_none=True)
        l = runner(*batch_args(d))
        (l[0] + l[1] + l[2]).sum().backward()
        return l

    W = max(3, args.warmup)
    ms_step, launches, clocks = H.timed(lambda: step(resident), args.steps, W)
    value = B * world / (ms_step / 1e3)

    # ---- e2e: the loader-side public API feeds the model from pinned host memory, every step ------------------------
    e2e = None
    if not args.no_e2e:
        # the same box counts as the resident batch, so that both legs run the same number of valid rows
        raw = raw_region_batch(B, V - 1, seed=4321 + rank, full=True)
        nb = (host["image_mask"].sum(1) - 1).to(torch.int32)
        live = torch.arange(V - 1).unsqueeze(0) < nb.unsqueeze(1)
        raw["raw_feat"] = raw["raw_feat"] * live.unsqueeze(-1)
        raw["raw_box"] = raw["raw_box"] * live.unsqueeze(-1)
        raw["sel"] = raw["sel"] * live.to(raw["sel"].dtype)
        raw["num_boxes"] = nb
        if args.raw_input == "bf16":
            raw["raw_feat"] = raw["raw_feat"].to(torch.bfloat16)
        hb = {k: v for k, v in host.items() if k not in ("image_feat", "image_loc", "image_mask", "image_label")}
        hb.update(raw)
        hb = {k: v.pin_memory() for k, v in hb.items()}

        def forever():
            while True:
                yield hb
        asm = RegionBatchAssembler(V - 1, 2048, cfg.num_locs, cfg.add_global_imgfeat,
                                   out_dtype=torch.bfloat16 if args.precision == "bf16" else torch.float32)
        loader = PrefetchingRegionLoader(forever(), asm, dev, depth=2)
        it = iter(loader)
        losses_host = torch.empty(3, dtype=torch.float32).pin_memory()
        loss_ready = torch.cuda.Event()

        def step_e2e():
            d = next(it)
            model.zero_grad(set_to_none=True)
            l = runner(*batch_args(d))
            # the losses are final once the forward has run: their device->host copy is queued before the backward, so
            # the host reads them (every step, as train_concap.py:316-318 does) while the backward is still executing
            losses_host.copy_(torch.cat([x.detach().reshape(1) for x in l]), non_blocking=True)
            loss_ready.record()
            (l[0] + l[1] + l[2]).sum().backward()
            loss_ready.synchronize()                   # the caller holds this step's three losses from here on
            return l
        ms_e2e, _, _ = H.timed(step_e2e, args.steps, 3)
        e2e = {"value": B * world / (ms_e2e / 1e3), "unit": "samples/s", "h2d_bytes_per_step": loader.h2d_bytes, "d2h_bytes_per_step": 12,
               "ms_per_step": ms_e2e,
               "input": "raw detector output (%s features, pixel boxes, image sizes, masking draws) + text tensors from pinned host "
                        "memory through volta_b200.pipeline.PrefetchingRegionLoader; batch assembled on the device by "
                        "vb_region_batch" % args.raw_input}
        it.close()

    # ---- the same step on a batch without padding (every len = T, every region valid): tensor-core use without the
    #      bonus of skipped rows.  Same graphs: the row counts live on the device.
    full = None
    if not args.no_full_length_leg and not args.full_length and args.precision == "bf16":
        fl = {k: v.to(dev) for k, v in pretraining_batch(B, T, V, seed=1234 + rank, full_length=True).items()}
        ms_full, _, _ = H.timed(lambda: step(fl), args.steps, 2)
        full = {"value": B * world / (ms_full / 1e3), "unit": "samples/s", "ms_per_step": ms_full,
                "what": "same step, batch without padding (every text length = T, every region valid)"}
        key = (args.model, T, V)
        if key in FLOPS_PER_SAMPLE:
            full["whole_step_frac_algorithmic"] = FLOPS_PER_SAMPLE[key] * 1e9 * (full["value"] / world) / (peaks()["burst"] * 1e12)

    # ---- optimizer-included step (reported next to the headline, which is the path alone): the
    #      reference's clip_grad_norm_(5.0) + AdamW with one group per parameter + zero_grad, through FusedAdamW
    with_opt = None
    if not args.no_optimizer_leg:
        from volta_b200.optim import FusedAdamW
        no_decay = ("bias", "LayerNorm.bias", "LayerNorm.weight")
        groups = [{"params": [p], "lr": 1e-4, "weight_decay": 0.0 if any(nd in n for nd in no_decay) else 0.01}
                  for n, p in model.named_parameters()]
        opt = FusedAdamW(groups, lr=1e-4, betas=(0.9, 0.999), eps=1e-6, model=model, max_grad_norm=5.0)

        def step_opt():
            opt.zero_grad()
            l = runner(*batch_args(resident))
            (l[0] + l[1] + l[2]).sum().backward()
            opt.step()
        ms_opt, _, _ = H.timed(step_opt, args.steps, 2)
        with_opt = {"value": B * world / (ms_opt / 1e3), "unit": "samples/s", "ms_per_step": ms_opt,
                    "what": "fwd + bwd + clip_grad_norm(5.0) + AdamW (one group per parameter) + zero_grad"}
        del opt

    def eager_step():
        model._use_cuda_graphs = False
        step(resident)
        model._use_cuda_graphs = True
    roofline = H.gemm_roofline(eager_step, ms_step, clocks, B, FLOPS_PER_SAMPLE.get((args.model, T, V)), "sample")
    tpath = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
    if (args.model, T, V, B) == ("ctrl_vilbert_base", 36, 37, 256) and os.path.exists(tpath) and not args.full_length:
        roofline["traffic"] = json.load(open(tpath)).get("dram_bytes_per_step")
        roofline["traffic_note"] = "DRAM read+write bytes of all GEMM launches of one step (ncu --set full, profiles/r02_gemm_traffic.json)"

    line = {"metric": metric_name(args), "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline}
    if e2e is not None:
        line["e2e"] = e2e
    if full is not None:
        line["full_length"] = full
    if with_opt is not None:
        line["with_optimizer"] = with_opt
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_reference(args, 3, 1)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class _Owned:
    """Presents a stream of THIS rank's work items to evaluate_retrieval so that its ownership rule (index % world == rank)
    selects every one of them: item k is yielded at global index k * world + rank (the other indices are placeholders the
    loop skips before touching them)."""

    def __init__(self, items, world, rank):
        self.items, self.world, self.rank = items, world, rank

    def __iter__(self):
        for b in self.items:
            for r in range(self.world):
                yield b if r == self.rank else None


# ------------------------------------------------------------------------------------------ GPU arm: retrieval scoring
def main_retrieval(args):
    """BASELINE.json configs[4]: the scoring loop of eval_retrieval.py:168-199 - one caption against `pairs` candidate
    images per model call, forward only - through BertForVLTasks and volta_b200.retrieval.evaluate_retrieval.  Every rank
    owns the work items `step % world == rank` (replicas, no data-path collective); the score matrices are merged once."""
    import contextlib
    H = Harness(args)
    torch, dist, dev, world, rank = H.torch, H.dist, H.dev, H.world, H.rank
    import volta_b200
    from volta_b200.retrieval import RetrievalRanker, evaluate_retrieval

    cfg = make_cfg(args)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(sys.stderr):
        model = volta_b200.BertForVLTasks(cfg, retrieval_task_cfg(), [RETRIEVAL_TASK]).set_precision(args.precision)
    model.eval().to(dev)
    P = args.pairs
    n_items = 4
    ranker = RetrievalRanker(num_captions=n_items * world, num_images=P, device=dev, chunk=P)
    # every rank's own work items: global item index g = i * world + rank (so that owns(g) holds on this rank)
    host_items = [retrieval_item(args, 5000 + 7 * (i * world + rank), step=i * world + rank) for i in range(n_items)]
    # the two index fields stay host integers, as the reference's loader yields them
    resident = [tuple(t.to(dev) for t in it_[:7]) + it_[7:] for it_ in host_items]
    # e2e: the region features travel in --raw-input's dtype (bf16 by default: what the first GEMM reads; fp32 = as stored)
    wire = torch.bfloat16 if args.raw_input == "bf16" and args.precision == "bf16" else torch.float32
    pinned = [tuple((t.to(wire) if k == 0 else t).pin_memory() for k, t in enumerate(it_[:7])) + it_[7:] for it_ in host_items]
    h2d = sum(t.numel() * t.element_size() for t in pinned[0][:7])
    state = {"i": 0}

    def one(items):
        i = state["i"] % n_items
        state["i"] += 1
        # global index of this rank's i-th item is i * world + rank: evaluate_retrieval's ownership rule (index % world ==
        # rank) selects it on this rank and nowhere else
        evaluate_retrieval(model, RETRIEVAL_TASK, [items[i]], ranker, dev, start=i * world + rank, merge=False)

    W = max(3, args.warmup)
    ms_step, launches, clocks = H.timed(lambda: one(resident), args.steps, W)
    value = P * world / (ms_step / 1e3)
    e2e = None
    if not args.no_e2e:
        scores_host = torch.empty(P, dtype=torch.float32).pin_memory()

        # one evaluate_retrieval call over the whole timed stream of items: its prefetcher uploads item i+1 on a copy
        # stream while item i is scored; after every item the host reads that item's 500 scores back
        def e2e_items(n):
            for k in range(n):
                yield pinned[k % n_items]

        plain_update = ranker.update

        def update_and_read_back(c, i, logit, target):     # the step's result: its 500 scores, read by the host
            plain_update(c, i, logit, target)
            scores_host.copy_(ranker.score_matrix[int(c), :P], non_blocking=True)

        def run_e2e(n):
            # every rank scores all of ITS items: a one-rank ownership view of the stream
            ranker.update = update_and_read_back
            try:
                evaluate_retrieval(model, RETRIEVAL_TASK, _Owned(e2e_items(n), world, rank), ranker, dev, merge=False)
            finally:
                ranker.update = plain_update
            torch.cuda.current_stream().synchronize()
        run_e2e(3)
        H.barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(args.steps)
        e1.record()
        H.barrier()
        ms_e2e = e0.elapsed_time(e1) / args.steps
        if world > 1:
            tt = torch.tensor([ms_e2e], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_e2e = float(tt)
        e2e = {"value": P * world / (ms_e2e / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * P,
               "ms_per_step": ms_e2e, "input": "one work item (%s features) from pinned host memory per step, uploaded on a copy stream while the previous "
                                                 "item is scored; the item's 500 scores are read back" % args.raw_input}
    ranker.merge()

    def eager_item():          # the instrumented pass brackets every GEMM launch: no graph replay here
        prev = os.environ.get("VOLTA_B200_GRAPHS")
        os.environ["VOLTA_B200_GRAPHS"] = "0"
        try:
            one(resident)
        finally:
            if prev is None:
                del os.environ["VOLTA_B200_GRAPHS"]
            else:
                os.environ["VOLTA_B200_GRAPHS"] = prev
    roofline = H.gemm_roofline(eager_item, ms_step, clocks, P, FLOPS_PER_PAIR.get((args.model, args.seq_len, args.regions)), "pair")
    line = {"metric": metric_name(args), "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
            "data": "synthetic", "config": workload_config(args, world), "clocks": clocks, "gpu_launches": launches,
            "roofline": roofline}
    if e2e is not None:
        line["e2e"] = e2e
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_reference(args, 2, 1)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def _claim_stdout():
    """stdout carries exactly ONE line, the JSON line: whatever libraries write to file descriptor 1 (NCCL prints its
    version banner there) goes to stderr instead; emit() writes the line to the real stdout."""
    global _JSON_FD
    try:
        sys.stdout.flush()
        fd = os.dup(1)
        os.dup2(2, 1)
        _JSON_FD = fd
    except OSError:       # no usable stderr / stdout descriptor: print the line the ordinary way
        _JSON_FD = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    args = parse()
    _claim_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "retrieval":
        return main_retrieval(args)
    return main_pretrain(args)


if __name__ == "__main__":
    main()
