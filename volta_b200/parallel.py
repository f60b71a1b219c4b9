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


class SwitchAllReduce:
    """The flat fp32 gradient buffer in SYMMETRIC memory (the same allocation on every rank, peer-mapped and bound to an
    NVSwitch multicast address by torch.distributed._symmetric_memory - plumbing), reduced by this library's own kernel
    `vb_allreduce_f32` (csrc/allreduce.cu): in-switch reduction with multimem.ld_reduce / multimem.st, no shared memory,
    so its blocks are co-resident with the backward's persistent GEMM CTAs instead of displacing them.  One launch per
    bucket on a high-priority communication stream; `wait()` joins it into the current stream."""

    def __init__(self, numel, device, group=None, blocks=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        pg = group if group is not None else dist.distributed_c10d._get_default_group()
        need = int(_lib.load().vb_allreduce_pad_bytes())
        if symm_mem.get_signal_pad_size() < need:
            symm_mem.set_signal_pad_size(need)
        self.buf = symm_mem.empty(numel, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, pg)
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        self.multicast = int(self.hdl.multicast_ptr or 0)
        if os.environ.get("VOLTA_B200_DDP_MULTICAST", "1") == "0":
            self.multicast = 0
        # the tensor may sit at an offset inside its (symmetric) allocation: the same offset on every rank
        delta = self.buf.data_ptr() - int(self.hdl.buffer_ptrs[self.rank])
        assert 0 <= delta < (1 << 40), "unexpected symmetric-memory layout"
        self._ptr_table = torch.tensor([int(p) + delta for p in self.hdl.buffer_ptrs], dtype=torch.int64, device=device)
        self.buffer_ptrs_dev = self._ptr_table.data_ptr()
        if self.multicast:
            self.multicast += delta
        self.signal_pad_ptrs_dev = int(self.hdl.signal_pad_ptrs_dev)
        self.blocks = int(blocks if blocks is not None else os.environ.get("VOLTA_B200_DDP_BLOCKS", "64"))
        self.blocks_alone = int(os.environ.get("VOLTA_B200_DDP_BLOCKS_ALONE", "296"))
        self.stream = torch.cuda.Stream(device=device, priority=-1)
        self.device = device
        self.launches = 0
        # our flag region of the signal pad starts zeroed on every rank before anyone launches
        pad = self.hdl.get_signal_pad(self.rank, (need // 4,), dtype=torch.int32)
        pad[1024:].zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group=pg)

    @property
    def kind(self):
        return "nvswitch multimem" if self.multicast else "peer pointers (two-shot)"

    def allreduce(self, lo, hi, scale, alone=False):
        """alone=True: nothing else will run next to this launch (the last bucket of a step) - two blocks per SM, twice the
        loads in flight; every rank must make the same choice."""
        from . import _lib
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)                       # the bucket's gradients are final in stream order
        _lib.call("vb_allreduce_f32", self.buffer_ptrs_dev, self.multicast, self.signal_pad_ptrs_dev, self.rank, self.world,
                  int(lo), int(hi - lo), float(scale), self.blocks_alone if alone else self.blocks, self.stream.cuda_stream)
        self.launches += 1

    def wait(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)


class GradBucketer:
    """Slices a flat gradient buffer into buckets that close as the backward frontier moves towards offset 0.

    `boundaries`: ascending element offsets at which a bucket may end (sub-layer boundaries of the model).
    `advance(frontier)` is called with the lowest offset whose gradient is final; every closed bucket of at
    least `message_size` elements is all-reduced asynchronously.  `finish()` flushes the rest and waits.
    Pure host logic + torch.distributed: works with gloo on CPU tensors (tests) and NCCL on CUDA."""

    def __init__(self, flat_grad, boundaries, message_size=10_000_000, group=None, average=True, wire_dtype=None, wire_buffer=None,
                 reducer=None):
        self.flat = flat_grad
        self.reducer = reducer    # SwitchAllReduce over the same buffer: buckets go through vb_allreduce_f32 instead of NCCL
        # Optional reduced-precision wire format (SURVEY.md 8e "bf16 or fp32 buckets"; apex reduces half-precision
        # gradients in their own dtype unless allreduce_always_fp32 is set, distributed.py:442-456): the bucket is cast
        # into `wire` when it closes, the collective moves half the bytes, the result is cast back in finish().
        self.wire = None
        if wire_dtype is not None and wire_dtype != flat_grad.dtype:
            self.wire = wire_buffer if wire_buffer is not None else torch.empty(flat_grad.numel(), dtype=wire_dtype, device=flat_grad.device)
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
        if self.reducer is not None:
            self.reducer.allreduce(lo, hi, 1.0 / self.world if self.average else 1.0, alone=(lo == 0))
            self.works.append((None, None, None))
            self.ranges.append((lo, hi))
            self.hi = lo
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
        if self.reducer is not None:
            self.reducer.wait()
            n = len(self.works)
            self.works = []
            return n
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
        self._reducer = self._make_switch_allreduce()
        self.registered = False if self._reducer is not None else self._register_gradient_buffer()
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

    def _make_switch_allreduce(self):
        """VOLTA_B200_DDP_ALLREDUCE = "switch" (default on CUDA with the fp32 wire): the gradient buffer lives in symmetric
        memory and buckets are reduced by vb_allreduce_f32; "nccl": torch.distributed all_reduce.  Falls back to NCCL
        (with a warning) when symmetric memory cannot be set up."""
        mode = os.environ.get("VOLTA_B200_DDP_ALLREDUCE", "switch")
        if (self.world_size == 1 or mode != "switch" or self.allreduce_dtype not in (None, torch.float32) or
                not torch.cuda.is_available() or dist.get_backend(self.group) != "nccl"):
            return None
        store = engine.model_store(self.module)
        if store.grad32 is not None or not store.flat32.is_cuda:
            return None
        try:
            red = SwitchAllReduce(store.total, store.flat32.device, self.group)
        except Exception as e:
            import warnings
            warnings.warn("volta_b200 DDP: symmetric-memory all-reduce unavailable (%r); using NCCL" % (e,))
            return None
        store.grad32 = red.buf
        return red

    def _register_gradient_buffer(self):
        """NCCL user-buffer registration of the flat gradient buffer (one registration for every bucket, since a bucket is a
        slice of it): the buffer is allocated by NCCL's own allocator (ncclMemAlloc through torch's MemPool) and the pool is
        registered with the communicator, so the NVLink-SHARP (NVLS) all-reduce reads and writes it in place - no staging
        copies through NCCL's internal buffers, fewer SMs taken from the backward GEMMs.  VOLTA_B200_DDP_REGISTER=0 keeps
        a plain allocation.  Returns True when the registration took place."""
        if self.world_size == 1 or os.environ.get("VOLTA_B200_DDP_REGISTER", "1") == "0" or not torch.cuda.is_available():
            return False
        store = engine.model_store(self.module)
        if store.grad32 is not None or not store.flat32.is_cuda or dist.get_backend(self.group) != "nccl":
            return False
        try:
            pg = self.group if self.group is not None else dist.distributed_c10d._get_default_group()
            backend = pg._get_backend(store.flat32.device)
            pool = torch.cuda.MemPool(backend.mem_allocator)
            with torch.cuda.use_mem_pool(pool):
                grad32 = torch.empty(store.total, dtype=torch.float32, device=store.flat32.device)
                wire = None
                if self.allreduce_dtype not in (None, torch.float32):
                    wire = torch.empty(store.total, dtype=self.allreduce_dtype, device=store.flat32.device)
            backend.register_mem_pool(pool)
        except Exception as e:  # older torch / NCCL: plain buffers work, only slower
            import warnings
            warnings.warn("volta_b200 DDP: NCCL buffer registration unavailable (%s); using unregistered buffers" % (e,))
            return False
        store.grad32 = grad32
        self._pool, self._wire = pool, wire   # the pool must outlive the buffers
        return True

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
                                          wire_dtype=self.allreduce_dtype, wire_buffer=getattr(self, "_wire", None),
                                          reducer=self._reducer if (self._reducer is not None and
                                                                    self._reducer.buf.data_ptr() == store.grad32.data_ptr()) else None)
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
