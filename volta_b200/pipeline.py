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

    def __call__(self, raw_feat, raw_box, image_hw, num_boxes, sel, device=None):
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
        V = self.num_rows
        feat = torch.empty(B, V, F, dtype=self.out_dtype, device=device)
        loc = torch.empty(B, V, self.num_locs, dtype=torch.float32, device=device)
        mask = torch.empty(B, V, dtype=torch.int64, device=device)
        label = torch.empty(B, R, dtype=torch.int64, device=device)
        count = torch.empty(B, dtype=torch.int32, device=device)
        if B == 0:        # an empty batch has nothing to assemble (and no buffers to point at)
            return feat, loc, mask, label, count
        with torch.cuda.device(device):
            ops.region_batch(raw_feat, raw_box, image_hw, num_boxes, sel, self.global_pos, feat, loc, mask, label, count)
        return feat, loc, mask, label, count
