"""CPU oracle for the region batch assembly of the Conceptual-Captions loader.  TEST INFRASTRUCTURE ONLY.

numpy restatement (fp32, the reference's operation order) of what the reference does on the host for every batch
between the detector output and the model call:
  * box area and normalisation            volta/datasets/concept_cap_dataset.py:433-460 (BertPreprocessBatch.__call__)
  * box-to-box IoU                        volta/datasets/concept_cap_dataset.py:31-68   (iou)
  * region masking + overlap exclusion    volta/datasets/concept_cap_dataset.py:661-690 (random_region)
  * padding masks / labels                volta/datasets/concept_cap_dataset.py:562-568
  * global mean feature, box and mask     volta/datasets/concept_cap_dataset.py:237-264 (ConceptCapLoaderTrain.__iter__)
The random draws of `random_region` are an INPUT here (`sel`: 0 = untouched, 1 = predicted and zeroed, 2 = predicted,
features kept), so the function is deterministic.

Parity status: PINNED.  `tests/golden/make_golden_pipeline.py` runs the UNMODIFIED reference classes
(`BertPreprocessBatch.__call__`, `ConceptCapLoaderTrain.__iter__`; stubs only for the absent `tensorpack` /
`msgpack_numpy` imports, a toy tokenizer) on seeded detector outputs and commits their outputs under
`tests/golden/pipeline_*.pt`; `tests/test_pipeline.py` checks this file against them BIT FOR BIT.

Nothing in `volta_b200/` imports this module; only `tests/` (incl. the timed CPU-baseline leg of `tests/perf/`) and
`__graft_entry__.smoke()` may.
"""
import numpy as np

f32 = np.float32


def iou(boxes):
    """[n,4] pixel boxes -> [n,n] overlaps, fp32, the reference's operation order (concept_cap_dataset.py:31-68)."""
    b = boxes.astype(f32)
    area = ((b[:, 2] - b[:, 0] + f32(1)) * (b[:, 3] - b[:, 1] + f32(1))).astype(f32)
    iw = np.minimum(b[:, None, 2], b[None, :, 2]) - np.maximum(b[:, None, 0], b[None, :, 0]) + f32(1)
    iw = np.where(iw < 0, f32(0), iw).astype(f32)
    ih = np.minimum(b[:, None, 3], b[None, :, 3]) - np.maximum(b[:, None, 1], b[None, :, 1]) + f32(1)
    ih = np.where(ih < 0, f32(0), ih).astype(f32)
    inter = (iw * ih).astype(f32)
    ua = ((area[:, None] + area[None, :]).astype(f32) - inter).astype(f32)
    return (inter / ua).astype(f32)


def assemble_region_batch(raw_feat, raw_box, image_hw, num_boxes, sel, num_locs=5, add_global_imgfeat="first"):
    """raw_feat [B,R,F] fp32, raw_box [B,R,4] fp32 pixels, image_hw [B,2] (h, w), num_boxes [B], sel [B,R] uint8.
    Returns image_feat [B,V,F] fp32, image_loc [B,V,num_locs] fp32, image_mask [B,V] int64, image_label [B,R] int64,
    sum_count [B] int64 (the divisor of the global feature)."""
    raw_feat = np.asarray(raw_feat, dtype=f32)
    raw_box = np.asarray(raw_box, dtype=f32)
    B, R, F = raw_feat.shape
    feat = np.zeros((B, R, F), f32)
    loc = np.zeros((B, R, num_locs), f32)
    label = np.full((B, R), -1, np.int64)
    mask = np.zeros((B, R), np.int64)
    count = np.zeros((B,), np.int64)
    for b in range(B):
        nb = int(num_boxes[b])
        h, w = f32(image_hw[b][0]), f32(image_hw[b][1])
        feat[b, :nb] = raw_feat[b, :nb]                               # :441 (rows beyond the count stay zero)
        loc[b, :nb, :4] = raw_box[b, :nb]                             # :444
        if num_locs == 5:                                             # :450-455, area of the PIXEL box over w*h
            wh = f32(float(w) * float(h))
            loc[b, :, 4] = ((loc[b, :, 3] - loc[b, :, 1]) * (loc[b, :, 2] - loc[b, :, 0])).astype(f32) / wh
        loc[b, :, 0] = loc[b, :, 0] / w                               # :458-461
        loc[b, :, 1] = loc[b, :, 1] / h
        loc[b, :, 2] = loc[b, :, 2] / w
        loc[b, :, 3] = loc[b, :, 3] / h
        ov = iou(raw_box[b, :nb])                                     # :438
        masked = np.zeros((R,), bool)                                 # :665
        for i in range(nb):                                           # :667-688
            if sel[b][i] != 0:
                if sel[b][i] == 1:
                    feat[b, i] = 0                                    # :676
                masked[:nb] |= ov[i] > f32(0.4)                       # :679
                label[b, i] = 1
        mask[b, :nb] = 1                                              # :562-566
        c = int((~masked).sum())                                      # :238 counts the padded rows too
        count[b] = c if c != 0 else 1                                 # :239
    # :240 np.sum over the region axis accumulates row after row in fp32; the division by an int64 count happens in
    # fp64 and the result is cast back to fp32 (:242) - one correctly rounded fp32 division
    total = np.zeros((B, F), f32)
    for r in range(R):
        total = (total + feat[:, r]).astype(f32)
    g_feat = (total.astype(np.float64) / count[:, None]).astype(f32)
    if add_global_imgfeat in ("first", "last"):
        g_loc = np.array([0, 0, 1, 1] + [1] * (num_locs - 4), f32)    # :244
        g_loc = np.repeat(g_loc[None, None], B, 0)
        ones = np.ones((B, 1), np.int64)
        if add_global_imgfeat == "first":
            feat = np.concatenate([g_feat[:, None], feat], 1)
            loc = np.concatenate([g_loc, loc], 1)
            mask = np.concatenate([ones, mask], 1)
        else:                                                         # :252-264
            feat = np.concatenate([feat, g_feat[:, None]], 1)
            loc = np.concatenate([loc, g_loc], 1)
            mask = np.concatenate([mask, ones], 1)
    return feat, loc, mask, label, count
