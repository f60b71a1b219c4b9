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
  e2e       the same metric through the module API with HOST (pinned) buffers: every step copies its batch
            host->device (on a copy stream, while the previous step computes) and reads its three losses back; the
            losses are final after the forward, so their read-back is queued before the backward and the host blocks
            on it - not on the backward - before it queues the next step (all work is inside the timed region)
  roofline  dominant kernel = the tcgen05 GEMM: algorithmic FLOPs (2*M*N*K per launch) / CUDA-event time of
            every GEMM launch in an instrumented step, against the measured sustained bf16 peak
  cpu_baseline / --impl reference: the oracle port of the reference (oracle/volta_oracle.py, fp32, train mode)
            timed on this box's host cores on a bounded sample (batch 8 per step)
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

METRIC = "ctrl_vilbert pretrain samples/s"
FLOPS_PER_SAMPLE = {  # algorithmic fwd+bwd GFLOP per sample of the REFERENCE algorithm
    ("ctrl_vilbert_base", 36, 37): 52.578, ("ctrl_visualbert_base", 36, 37): 43.627, ("ctrl_lxmert", 36, 37): 48.537,
    ("ctrl_uniter_base", 36, 37): 43.628, ("vilbert_base", 30, 101): 77.646,
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="volta_b200", choices=["volta_b200", "reference"])
    ap.add_argument("--model", default="ctrl_vilbert_base")
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--seq-len", type=int, default=36)
    ap.add_argument("--regions", type=int, default=37, help="visual tokens including the global feature")
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--raw-input", default="", choices=["", "fp32", "bf16"],
                    help="e2e leg: upload the RAW detector output (features in this dtype, pixel boxes, image sizes, the "
                         "masking draws) and assemble image_feat / image_loc / image_mask / image_label on the device "
                         "(volta_b200.pipeline.RegionBatchAssembler) inside the timed step")
    ap.add_argument("--no-optimizer-leg", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(tflops=d.get("bf16_tflops_sustained", 1354.0), burst=d.get("bf16_tflops", 1651.7),
                    hbm=d.get("hbm_gbs", 6529.7), source="measured (MEASURED_PEAKS.json, sustained)")
    return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


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
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
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


def workload_config(args, world):
    B, T, V = args.batch, args.seq_len, args.regions
    return {"workload": "%s pre-training fwd+bwd (MLM+MRM+ITM, train mode, dropout on), per-GPU batch %d, "
                        "T=%d text tokens, V=%d regions incl. global x 2048-d + 5-d boxes" % (args.model, B, T, V),
            "per_gpu_batch": B, "global_batch": B * world, "parallelism": "dp%d" % world,
            "l2": "working set (activations ~25 MB/sample, bf16 weights 484 MB) exceeds the 126 MB L2; no flush needed",
            "launch": "CUDA graphs (forward graph + backward graph segments), replayed per step",
            "mlm_head": "vocabulary decoder + cross-entropy run on the labelled (masked) rows only, count kept on the device; "
                        "the reference computes all T positions and ignores ~90 % of them (ignore_index=-1) - same loss and "
                        "gradients.  whole_step_frac_of_burst_peak uses the reference algorithm's FLOPs; the executed GEMM "
                        "FLOPs are in roofline.executed_gemm_gflop_per_sample (VOLTA_B200_MLM_COMPACT=0 runs all positions)"}


def make_cfg(args, dropout=True):
    from volta_b200.config import named_config
    over = {} if dropout else dict(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, v_hidden_dropout_prob=0.0,
                                   v_attention_probs_dropout_prob=0.0)
    return named_config(args.model, **over)


def batch_args(d):
    return (d["input_ids"], d["image_feat"], d["image_loc"], d["segment_ids"], d["input_mask"], d["image_mask"],
            d["lm_label_ids"], d["image_label"], d["image_cls"], None, None, None, None, None, d["is_match"])


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference(args, steps, warmup, quiet=False):
    """The reference algorithm (oracle port, fp32, train mode with dropout) on the host cores."""
    import torch
    from oracle import volta_oracle as O
    import volta_b200
    from volta_b200.synthetic import pretraining_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = make_cfg(args)
    torch.manual_seed(0)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        model = volta_b200.BertForVLPreTraining(cfg)  # parameter container only: supplies reference-initialised weights
    P = O.params_from_state_dict(model.state_dict(), requires_grad=True)
    B = args.cpu_batch
    b = pretraining_batch(B, args.seq_len, args.regions, seed=1234)
    times = []
    for i in range(warmup + steps):
        for v in {id(v): v for v in P.values()}.values():
            v.grad = None
        t0 = time.perf_counter()
        O.fwd_bwd(P, cfg, b, train=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return dict(value=B / med, unit="samples/s", cores=torch.get_num_threads(), kind="port",
                sample="%s fp32 train-mode fwd+bwd, batch %d, T=%d V=%d, %d warm-up + %d timed steps, median %.3f s/step"
                       % (args.model, B, args.seq_len, args.regions, warmup, steps, med)), med


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    cb, med = cpu_reference(args, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "samples/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {k: v for k, v in workload_config(args, max(1, args.gpus)).items() if k not in ("launch", "l2", "mlm_head")},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import volta_b200
    from volta_b200 import engine
    from volta_b200.synthetic import pretraining_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: volta_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("VOLTA_B200_NCCL_CTAS"):  # experiments with VOLTA_B200_DDP_RESERVE_SMS (parallel.py)
            os.environ.setdefault("NCCL_MAX_CTAS", os.environ["VOLTA_B200_NCCL_CTAS"])
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner must not share stdout with the JSON line
        dist.init_process_group(backend="nccl", device_id=dev)

    cfg = make_cfg(args)
    torch.manual_seed(0)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):  # the constructor prints the visual targets, as the reference does
        model = volta_b200.BertForVLPreTraining(cfg).set_precision(args.precision)
    model.train().to(dev)
    runner = model
    if world > 1:
        from volta_b200.parallel import DistributedDataParallel
        runner = DistributedDataParallel(model)

    B, T, V = args.batch, args.seq_len, args.regions
    host = pretraining_batch(B, T, V, seed=1234 + rank)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    losses_host = torch.empty(3, dtype=torch.float32).pin_memory()

    def step(d):
        model.zero_grad(set_to_none=True)
        l = runner(*batch_args(d))
        (l[0] + l[1] + l[2]).sum().backward()
        return l

    # e2e: every step uploads ITS batch from pinned host memory and reads the three losses back.  The upload of
    # batch i+1 is issued on a copy stream while step i computes (what a prefetching loader does; the reference's
    # loop does `batch.cuda(non_blocking=True)`, train_concap.py:273); the step waits on the upload's event.
    copy_stream = torch.cuda.Stream(device=dev)
    loss_ready = torch.cuda.Event()
    pending = []

    def upload():
        main = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            d = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        for v in d.values():
            v.record_stream(main)
        pending.append((d, ev))

    asm = None
    if args.raw_input:
        # the loader-side leg of SURVEY.md 8f rank 2: what travels is the detector output as stored, the batch is assembled
        # on the device; image_feat stays fp32 so that the step replays the graph captured for the resident batch
        from volta_b200.pipeline import RegionBatchAssembler
        from volta_b200.synthetic import raw_region_batch
        raw = raw_region_batch(B, V - 1, seed=4321 + rank, full=False)
        if args.raw_input == "bf16":
            raw["raw_feat"] = raw["raw_feat"].to(torch.bfloat16)
        for k in ("image_feat", "image_loc", "image_mask", "image_label"):
            del pinned[k]
        pinned.update({k: v.pin_memory() for k, v in raw.items()})
        h2d = sum(v.numel() * v.element_size() for v in pinned.values())
        asm = RegionBatchAssembler(V - 1, 2048, 5, "first", out_dtype=torch.float32)

    def step_e2e():
        if not pending:
            upload()
        d, ev = pending.pop(0)
        torch.cuda.current_stream().wait_event(ev)
        if asm is not None:
            feat, loc, mask, label, _ = asm(d["raw_feat"], d["raw_box"], d["image_hw"], d["num_boxes"], d["sel"])
            d = dict(d, image_feat=feat, image_loc=loc, image_mask=mask, image_label=label)
        model.zero_grad(set_to_none=True)
        l = runner(*batch_args(d))
        # the losses are final once the forward has run: their device->host copy is queued before the backward, so the
        # host can read them (every step, as train_concap.py:316-318 does) while the backward is still executing
        losses_host.copy_(torch.cat([x.detach().reshape(1) for x in l]), non_blocking=True)
        loss_ready.record()
        (l[0] + l[1] + l[2]).sum().backward()
        upload()                                   # next step's batch travels while this step runs
        loss_ready.synchronize()                   # the caller holds this step's three losses from here on
        return l

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        n0 = engine.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as cs:
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps, engine.launch_count - n0, cs.summary()

    W = max(3, args.warmup)
    ms_step, launches, clocks = timed(lambda: step(resident), args.steps, W)
    value = B * world / (ms_step / 1e3)

    e2e = None
    if not args.no_e2e:
        ms_e2e, _, _ = timed(step_e2e, args.steps, 2)
        e2e = {"value": B * world / (ms_e2e / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
               "ms_per_step": ms_e2e}
        if asm is not None:
            e2e["input"] = "raw detector output (%s features), assembled on the device by vb_region_batch" % args.raw_input

    # optimizer-included step (reported next to the headline, which is the path alone): the
    # reference's clip_grad_norm_(5.0) + AdamW with one group per parameter + zero_grad, through volta_b200.optim.FusedAdamW
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
        ms_opt, _, _ = timed(step_opt, args.steps, 2)
        with_opt = {"value": B * world / (ms_opt / 1e3), "unit": "samples/s", "ms_per_step": ms_opt,
                    "what": "fwd + bwd + clip_grad_norm(5.0) + AdamW (one group per parameter) + zero_grad"}
        del opt

    # roofline leg: one instrumented EAGER step (the timed legs replay CUDA graphs, whose kernels cannot be bracketed
    # one by one).  Every GEMM launch of the step is re-issued 5x back to back between two CUDA events on the launching
    # stream (vb_gemm_bench) - device time per launch without host gaps, operands as L2-warm as in the real step.
    from volta_b200 import ops as vb_ops
    model._use_cuda_graphs = False
    vb_ops.GEMM_BENCH = []
    step(resident)
    torch.cuda.synchronize()
    recs, vb_ops.GEMM_BENCH = vb_ops.GEMM_BENCH, None
    model._use_cuda_graphs = True
    tc = [(f, ms) for f, ms, is_tc, _ in recs if is_tc]
    if os.environ.get("VOLTA_B200_BENCH_VERBOSE") and rank == 0:
        agg = {}
        for f, ms, is_tc, shape in recs:
            t = agg.setdefault((shape, is_tc), [0, 0.0, 0.0])
            t[0] += 1; t[1] += ms; t[2] += f
        for (shape, is_tc), (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print("gemm %-44s tc=%d n=%3d  %8.3f ms  %7.1f TFLOP/s" % (shape, is_tc, n, ms, fl / ms / 1e9 if ms > 0 else 0), file=sys.stderr)
    pk = peaks()
    flops_tc, ms_tc = sum(f for f, _ in tc), sum(t for _, t in tc)
    achieved = flops_tc / (ms_tc * 1e-3) / 1e12 if ms_tc > 0 else 0.0
    key = (args.model, T, V)
    model_frac = None
    if key in FLOPS_PER_SAMPLE:
        model_frac = FLOPS_PER_SAMPLE[key] * 1e9 * (value / world) / (pk["burst"] * 1e12)
    traffic = None  # DRAM bytes of the step's GEMM launches from the committed ncu capture of this exact workload
    tpath = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if key == ("ctrl_vilbert_base", 36, 37) and B == 256 and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_step")
    roofline = {"bound": "tensor", "kernel": "vb::tc::gemm_tc_kernel (tcgen05/TMEM/TMA)", "achieved": achieved,
                "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"], "traffic": traffic,
                "traffic_note": "DRAM read+write bytes of all GEMM launches of one step (ncu, profiles/r01_gemm_traffic.json); "
                                "achieved/peak are FLOP-weighted over the same launches",
                "peak_source": pk["source"], "launches_per_step": len(tc), "gemm_ms_per_step": ms_tc,
                "gemm_share_of_step": ms_tc / ms_step,
                "whole_step_frac_of_burst_peak": model_frac,
                "algorithmic_gflop_per_sample": FLOPS_PER_SAMPLE.get(key),
                "executed_gemm_gflop_per_sample": sum(f for f, _, _, _ in recs) / B / 1e9}

    line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline}
    if e2e is not None:
        line["e2e"] = e2e
    if with_opt is not None:
        line["with_optimizer"] = with_opt
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_reference(args, 3, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
