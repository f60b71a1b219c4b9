"""Input pipeline -> device: the per-batch assembly the reference's Conceptual-Captions loader
does on the host with numpy, done by one C-ABI call on the GPU.

Mirrors, for the region side of a batch,
  * `BertPreprocessBatch.__call__`   (volta/datasets/concept_cap_dataset.py:429-461): zero-padding to `region_len`, the
    5th box coordinate (area of the pixel box over w*h) and the normalisation of x by w and y by h;
  * `random_region`                  (:661-690): picked regions get label 1, 90 % of them are zeroed, every region with
    IoU > 0.4 to a picked one is left out of the global mean - the DRAWS stay on the host (`sel`), their effect on
    the tensors happens on the device;
  * `ConceptCapLoaderTrain.__iter__` (:237-264): the mean-pooled global feature, its box [0,0,1,1,1] and mask 1,
    first or last, per `config.add_global_imgfeat`.
The host uploads the detector output as it is stored (fp32 features, or bf16 to halve the copy; pixel boxes; image
sizes; box counts; one byte per region for the draws) from pinned memory; `image_feat` comes out in the dtype the
encoder's first GEMM reads (bf16 by default - the engine then skips its own cast pass), `image_loc` in fp32,
`image_mask` / `image_label` in int64 as the loader produces them.  In fp32 the outputs are bit-identical to the
reference's (tests/test_pipeline.py, tests/golden/pipeline_*.pt).

There is no CPU path: without a CUDA device the call raises.
"""
import queue
import threading

import torch

from . import ops

_GLOBAL = {None: 0, "none": 0, "": 0, False: 0, "first": 1, "last": 2}


class RegionBatchAssembler:
    """asm = RegionBatchAssembler(region_len, 2048, config.num_locs, config.add_global_imgfeat)
    image_feat, image_loc, image_mask, image_label, count = asm(raw_feat, raw_box, image_hw, num_boxes, sel)

    Inputs may be CPU tensors (pinned: copied asynchronously on the current stream) or CUDA tensors."""

    def __init__(self, region_len, feature_size=2048, num_locs=5, add_global_imgfeat="first", out_dtype=torch.bfloat16):
        if add_global_imgfeat not in _GLOBAL:
            raise ValueError("add_global_imgfeat must be 'first', 'last' or None, got %r" % (add_global_imgfeat,))
        if num_locs not in (4, 5):
            raise ValueError("num_locs must be 4 or 5 (concept_cap_dataset.py:450), got %r" % (num_locs,))
        if out_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("out_dtype must be float32 or bfloat16")
        self.region_len, self.feature_size, self.num_locs = int(region_len), int(feature_size), int(num_locs)
        self.global_pos = _GLOBAL[add_global_imgfeat]
        self.out_dtype = out_dtype

    @property
    def num_rows(self):
        return self.region_len + (1 if self.global_pos else 0)

    def _to_device(self, t, dtype, device):
        if t.dtype != dtype:
            t = t.to(dtype)
        return t.to(device, non_blocking=True).contiguous()

    def outputs(self, B, device):
        """A reusable set of output buffers for batches of B samples: (image_feat, image_loc, image_mask, image_label,
        count).  Passing them back as `out=` keeps the addresses fixed from call to call (CUDA-graph friendly, no
        allocator traffic in the loader loop)."""
        V, R, F = self.num_rows, self.region_len, self.feature_size
        return (torch.empty(B, V, F, dtype=self.out_dtype, device=device),
                torch.empty(B, V, self.num_locs, dtype=torch.float32, device=device),
                torch.empty(B, V, dtype=torch.int64, device=device),
                torch.empty(B, R, dtype=torch.int64, device=device),
                torch.empty(B, dtype=torch.int32, device=device))

    def __call__(self, raw_feat, raw_box, image_hw, num_boxes, sel, device=None, out=None):
        if device is None:
            device = raw_feat.device if raw_feat.is_cuda else (torch.device("cuda") if torch.cuda.is_available() else None)
        if device is None or torch.device(device).type != "cuda":
            raise RuntimeError("volta_b200.pipeline has no CPU path: a CUDA (sm_100a) device is required")
        B, R, F = raw_feat.shape
        if R != self.region_len or F != self.feature_size:
            raise ValueError("expected raw_feat [B,%d,%d], got %s" % (self.region_len, self.feature_size, tuple(raw_feat.shape)))
        in_dtype = raw_feat.dtype if raw_feat.dtype in (torch.float32, torch.bfloat16) else torch.float32
        raw_feat = self._to_device(raw_feat, in_dtype, device)
        raw_box = self._to_device(raw_box, torch.float32, device)
        image_hw = self._to_device(image_hw, torch.float32, device)
        num_boxes = self._to_device(num_boxes, torch.int32, device)
        sel = self._to_device(sel, torch.uint8, device)
        feat, loc, mask, label, count = self.outputs(B, device) if out is None else out
        if feat.shape[0] != B or feat.dtype != self.out_dtype or not feat.is_cuda:
            raise ValueError("out= buffers do not match this batch (use RegionBatchAssembler.outputs(B, device))")
        if B == 0:        # an empty batch has nothing to assemble (and no buffers to point at)
            return feat, loc, mask, label, count
        with torch.cuda.device(device):
            ops.region_batch(raw_feat, raw_box, image_hw, num_boxes, sel, self.global_pos, feat, loc, mask, label, count)
        return feat, loc, mask, label, count


class PrefetchingRegionLoader:
    """The device-facing half of the reference's training loader (volta/datasets/concept_cap_dataset.py:229-286: the
    tensorpack producer hands numpy batches to the loop, which calls `.cuda(non_blocking=True)` on each, train_concap.py:273).

        loader = PrefetchingRegionLoader(host_batches, RegionBatchAssembler(36, 2048, 5, "first"), device)
        for batch in loader:     # dict of CUDA tensors: image_feat / image_loc / image_mask / image_label + everything else
            losses = model(batch["input_ids"], batch["image_feat"], ...)

    `host_batches` yields dicts of CPU tensors: the assembler's inputs (raw_feat, raw_box, image_hw, num_boxes, sel) and any
    other per-batch tensor (input_ids, input_mask, segment_ids, lm_label_ids, image_cls, is_match ...), which are uploaded
    as they are.  A staging thread keeps `depth` slots in flight: for each host batch it (1) waits until the consumer has
    left the slot, (2) copies pageable tensors into the slot's pinned buffers (pinned tensors are used in place),
    (3) issues the host->device copies and the assembly kernel on a copy stream into the slot's fixed device buffers and
    records an event.  `__next__` only makes the consumer's stream wait on that event, so the upload and assembly of batch
    i+1 run while step i computes.  `wire_dtype=torch.bfloat16` rounds the region features on the host before they travel
    (half the bytes of the largest tensor); what the GEMM reads is bf16 either way."""

    RAW = ("raw_feat", "raw_box", "image_hw", "num_boxes", "sel")

    def __init__(self, host_batches, assembler, device, depth=2, wire_dtype=None):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("volta_b200.pipeline has no CPU path: a CUDA (sm_100a) device is required")
        if depth < 2:
            raise ValueError("depth must be >= 2 (one slot is being consumed while the next is staged)")
        if wire_dtype not in (None, torch.float32, torch.bfloat16):
            raise ValueError("wire_dtype must be None, torch.float32 or torch.bfloat16")
        self.batches, self.asm, self.device, self.depth, self.wire_dtype = host_batches, assembler, device, int(depth), wire_dtype
        self.copy_stream = torch.cuda.Stream(device=device)
        self.h2d_bytes = 0           # bytes uploaded for the last staged batch

    class _Slot:
        def __init__(self):
            self.pinned, self.dev, self.out = {}, {}, None
            self.ready = torch.cuda.Event()
            self.left = None         # event the consumer records when it has queued its last read of this slot

    def _stage(self, slot, hb):
        dev = self.device
        if slot.left is not None:
            slot.left.synchronize()
        nbytes = 0
        srcs = {}
        for k, t in hb.items():
            if k == "raw_feat" and self.wire_dtype is not None and t.dtype != self.wire_dtype:
                t = t.to(self.wire_dtype)
            if not t.is_pinned():
                buf = slot.pinned.get(k)
                if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                    buf = slot.pinned[k] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                buf.copy_(t)
                t = buf
            srcs[k] = t
            nbytes += t.numel() * t.element_size()
        self.h2d_bytes = nbytes
        with torch.cuda.stream(self.copy_stream):
            for k, t in srcs.items():
                d = slot.dev.get(k)
                if d is None or d.shape != t.shape or d.dtype != t.dtype:
                    d = slot.dev[k] = torch.empty(t.shape, dtype=t.dtype, device=dev)
                d.copy_(t, non_blocking=True)
            B = srcs["raw_feat"].shape[0]
            if slot.out is None or slot.out[0].shape[0] != B:
                slot.out = self.asm.outputs(B, dev)
            r = slot.dev
            self.asm(r["raw_feat"], r["raw_box"], r["image_hw"].float() if r["image_hw"].dtype != torch.float32 else r["image_hw"],
                     r["num_boxes"], r["sel"], device=dev, out=slot.out)
            slot.ready.record(self.copy_stream)

    def _producer(self, q, free, stop):
        try:
            with torch.cuda.device(self.device):
                for hb in self.batches:
                    slot = None
                    while slot is None and not stop.is_set():
                        try:
                            slot = free.get(timeout=0.1)
                        except queue.Empty:
                            pass
                    if stop.is_set():
                        return
                    self._stage(slot, hb)
                    q.put(slot)
            q.put(None)
        except BaseException as e:  # surfaced in the consumer
            q.put(e)

    def __iter__(self):
        q, free, stop = queue.Queue(), queue.Queue(), threading.Event()
        for _ in range(self.depth):
            free.put(self._Slot())
        th = threading.Thread(target=self._producer, args=(q, free, stop), daemon=True)
        th.start()
        prev = None
        try:
            while True:
                item = q.get()
                if prev is not None:       # the consumer has queued everything that reads the previous slot
                    prev.left = torch.cuda.Event()
                    prev.left.record(torch.cuda.current_stream(self.device))
                    free.put(prev)
                    prev = None
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                torch.cuda.current_stream(self.device).wait_event(item.ready)
                feat, loc, mask, label, count = item.out
                batch = {k: v for k, v in item.dev.items() if k not in self.RAW}
                batch.update(image_feat=feat, image_loc=loc, image_mask=mask, image_label=label, num_regions=count)
                prev = item
                yield batch
        finally:
            stop.set()
