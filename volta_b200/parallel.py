"""Data-parallel training across the GPUs of one box: one process per GPU, NCCL over NVLink.

Drop-in for the wrapper the reference uses, `apex.parallel.DistributedDataParallel`
(apex/apex/parallel/distributed.py:129-639; call site train_concap.py:241-253): parameters are broadcast
from rank 0 when the model is wrapped, and gradients are averaged across ranks during `backward()`.

What is different from apex, by design:
  * the gradients already live in ONE flat fp32 buffer (store.py), so a bucket is a slice of it - apex's
    flatten -> all_reduce -> unflatten copies (distributed.py:425-475, csrc/flatten_unflatten.cpp,
    multi_tensor_scale) do not exist here;
  * the bucket order is known from the launch plan (gradients complete from the heads back to the
    embeddings, i.e. from the END of the flat buffer to its start), so there is no first-iteration discovery
    pass (distributed.py:363-394) and no bucket-structure broadcast (:292-296);
  * the average is taken by NCCL (ReduceOp.AVG), not by a separate `mul_(1/world)` pass (:453-454).
The all-reduce of a bucket is issued as soon as the backward has retired every parameter in it and runs on
NCCL's stream while the remaining backward kernels run on the compute stream.
"""
import os

import torch
import torch.distributed as dist

from . import engine


def _convert(src, dst):
    """dst = src in dst's dtype (fp32 <-> bf16), on the current stream: vb_eltwise on the GPU; torch on CPU tensors
    (the gloo tests of the host logic)."""
    if src.is_cuda:
        from . import ops
        from ._lib import ELT_COPY
        ops.eltwise(ELT_COPY, src.view(1, -1), dst.view(1, -1))
    else:
        dst.copy_(src)


class GradBucketer:
    """Slices a flat gradient buffer into buckets that close as the backward frontier moves towards offset 0.

    `boundaries`: ascending element offsets at which a bucket may end (sub-layer boundaries of the model).
    `advance(frontier)` is called with the lowest offset whose gradient is final; every closed bucket of at
    least `message_size` elements is all-reduced asynchronously.  `finish()` flushes the rest and waits.
    Pure host logic + torch.distributed: works with gloo on CPU tensors (tests) and NCCL on CUDA."""

    def __init__(self, flat_grad, boundaries, message_size=10_000_000, group=None, average=True, wire_dtype=None):
        self.flat = flat_grad
        # Optional reduced-precision wire format (SURVEY.md 8e "bf16 or fp32 buckets"; apex reduces half-precision
        # gradients in their own dtype unless allreduce_always_fp32 is set, distributed.py:442-456): the bucket is cast
        # into `wire` when it closes, the collective moves half the bytes, the result is cast back in finish().
        self.wire = None
        if wire_dtype is not None and wire_dtype != flat_grad.dtype:
            self.wire = torch.empty(flat_grad.numel(), dtype=wire_dtype, device=flat_grad.device)
        self.total = flat_grad.numel()
        self.boundaries = sorted(set(b for b in boundaries if 0 < b < self.total))
        self.message_size = int(message_size)
        self.group = group
        self.average = average
        self.world = dist.get_world_size(group)
        backend = dist.get_backend(group)
        self._native_avg = average and backend == "nccl"
        self.reset()

    def reset(self):
        self.hi = self.total      # everything in [hi, total) has been handed to the collective
        self.works = []
        self.ranges = []          # issued (lo, hi) ranges, in issue order (for tests / logging)

    def _issue(self, lo, hi):
        if hi <= lo:
            return
        chunk = self.flat[lo:hi]
        sent = chunk
        if self.wire is not None:
            sent = self.wire[lo:hi]
            _convert(chunk, sent)
        op = dist.ReduceOp.AVG if self._native_avg else dist.ReduceOp.SUM
        work = dist.all_reduce(sent, op=op, group=self.group, async_op=True)
        self.works.append((work, chunk, sent))
        self.ranges.append((lo, hi))
        self.hi = lo

    def advance(self, frontier):
        """Gradients at offsets >= frontier are final."""
        frontier = max(0, min(int(frontier), self.hi))
        # close the largest bucket that ends at self.hi, starts on a boundary >= frontier and is big enough
        candidates = [b for b in self.boundaries if frontier <= b < self.hi]
        if frontier == 0:
            candidates.append(0)
        if not candidates:
            return
        lo = min(candidates)
        if self.hi - lo >= self.message_size or lo == 0:
            self._issue(lo, self.hi)

    def simulate(self):
        """cut(frontier) -> bool: would `advance(frontier)` issue a collective?  Replays the bucketing decisions
        on a private copy of the state; the CUDA-graph executor cuts the captured backward at those points."""
        state = {"hi": self.total}

        def cut(frontier):
            frontier = max(0, min(int(frontier), state["hi"]))
            candidates = [b for b in self.boundaries if frontier <= b < state["hi"]]
            if frontier == 0:
                candidates.append(0)
            if not candidates:
                return False
            lo = min(candidates)
            if (state["hi"] - lo >= self.message_size or lo == 0) and state["hi"] > lo:
                state["hi"] = lo
                return True
            return False
        return cut

    def finish(self):
        self._issue(0, self.hi)
        for work, chunk, sent in self.works:
            work.wait()
            if sent is not chunk:
                _convert(sent, chunk)
            if self.average and not self._native_avg:
                chunk.div_(self.world)
        n = len(self.works)
        self.works = []
        return n


class DistributedDataParallel(torch.nn.Module):
    """`DistributedDataParallel(model, delay_allreduce=False, message_size=1e7)` as in the reference's call
    (train_concap.py:253).  The wrapped model must be one of volta_b200's models on a CUDA device."""

    def __init__(self, module, message_size=10_000_000, delay_allreduce=False, group=None, allreduce_dtype=None, **unused):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (init_process_group(backend='nccl'))")
        self.module = module
        self.message_size = int(float(os.environ.get("VOLTA_B200_DDP_MESSAGE_SIZE", message_size)))
        self.delay_allreduce = bool(delay_allreduce)
        self.group = group
        self.world_size = dist.get_world_size(group)
        # allreduce_dtype=torch.bfloat16 (or VOLTA_B200_DDP_BF16=1): gradients cross NVLink as bf16 - half the 968 MB of
        # ctrl_vilbert per step; each element is rounded once before and the sum once per hop.  Off by default: the
        # fp32 wire is what the parity tests pin.
        if allreduce_dtype is None and os.environ.get("VOLTA_B200_DDP_BF16", "0") == "1":
            allreduce_dtype = torch.bfloat16
        if allreduce_dtype not in (None, torch.float32, torch.bfloat16):
            raise ValueError("allreduce_dtype must be None, torch.float32 or torch.bfloat16")
        self.allreduce_dtype = allreduce_dtype
        self._bucketer = None
        self.last_num_buckets = 0
        self._broadcast_params()
        engine.register_progress_hook(module, self._on_progress, self._make_cut_fn)
        # Optional: leave SMs to NCCL while the backward runs (the persistent GEMM assigns tiles to its blocks statically,
        # so a block displaced by an all-reduce block holds its tiles back).  Measured on 8 B200 (30 steps, B=256/GPU):
        # no reserve 21.29 ms/step; 16 SMs reserved 21.79 ms (NCCL_MAX_CTAS=16) / 23.34 ms (=8) - the GEMMs lose more
        # than the collectives gain, so it is off by default.
        reserve = int(os.environ.get("VOLTA_B200_DDP_RESERVE_SMS", "0"))
        if self.world_size > 1 and not self.delay_allreduce and reserve > 0 and torch.cuda.is_available():
            sms = torch.cuda.get_device_properties(self._device()).multi_processor_count
            engine.set_backward_sm_limit(module, sms - reserve)
        engine.register_grad_ready_hook(module, self._on_done)

    def _device(self):
        return next(self.module.parameters()).device

    def _broadcast_params(self):
        """One broadcast of the flat fp32 parameter buffer from rank 0 (apex: distributed.py:253)."""
        store = engine.model_store(self.module)
        store.ensure(self._device(), self.module.precision)
        dist.broadcast(store.flat32, 0, group=self.group)
        store.refresh()

    def _boundaries(self, store):
        """Offsets where a sub-layer's parameters start: legal bucket ends."""
        bounds = set()
        for m in self.module.modules():
            ps = list(m.parameters(recurse=False))
            for p in ps:
                bounds.add(store.offsets[id(p)][0])
        return sorted(bounds)

    def _ensure_bucketer(self):
        store = engine.model_store(self.module)
        if self._bucketer is None or self._bucketer.flat.data_ptr() != store.grad32.data_ptr():
            self._bucketer = GradBucketer(store.grad32, self._boundaries(store), self.message_size, self.group,
                                          wire_dtype=self.allreduce_dtype)
        return self._bucketer

    def _on_progress(self, frontier_offset):
        if self.delay_allreduce or self.world_size == 1:
            return
        b = self._ensure_bucketer()
        if b.hi == b.total and not b.works:
            b.reset()
        b.advance(frontier_offset)

    def _make_cut_fn(self):
        if self.delay_allreduce or self.world_size == 1:
            return lambda frontier: False
        return self._ensure_bucketer().simulate()

    def _on_done(self):
        if self.world_size == 1:
            return
        b = self._ensure_bucketer()
        self.last_num_buckets = b.finish()
        b.reset()

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)
