"""Seeded synthetic batches with the shapes and dtypes the reference's loaders produce.

Field list and dtypes follow the pre-training call in the reference
(`train_concap.py:275-289`) and the Conceptual-Captions loader
(`volta/datasets/concept_cap_dataset.py:229-286`): a mean-pooled global region
is prepended (row 0 of `image_feat`), 15 % of tokens / regions are masked,
`image_cls` is a 1601-way soft label per detected region.  All tensors are
created on the CPU from a `torch.Generator`, so the same seed gives the same
batch here, on the GPU box and for the CPU baseline.
"""
import torch


def pretraining_batch(batch, seq_len, num_regions, *, vocab_size=30522, feat_dim=2048, num_locs=5,
                      num_classes=1601, seed=1234, full_length=False, global_first=True, global_last=False,
                      detector_labels=False):
    """Returns a dict of CPU tensors.  `num_regions` counts the global row (V).
    global_last: the global feature is the LAST row (`add_global_imgfeat: "last"`, vl-bert_base.json).
    detector_labels: also returns hard object / attribute labels with confidences (the targets of xent_1600 / xent_400 /
    xent_1601, volta/losses.py:87-133), drawn from a separate generator so the other fields do not change."""
    if global_last:
        global_first = True
    g = torch.Generator().manual_seed(seed)
    B, T, V = batch, seq_len, num_regions
    n_det = V - 1 if global_first else V

    lens = torch.full((B,), T) if full_length else torch.randint(min(8, T), T + 1, (B,), generator=g)
    pos = torch.arange(T).unsqueeze(0)
    input_mask = (pos < lens.unsqueeze(1)).long()
    ids = torch.randint(min(1000, vocab_size // 2), vocab_size, (B, T), generator=g)
    ids[:, 0] = 101 % vocab_size
    ids[torch.arange(B), lens - 1] = 102 % vocab_size
    ids = ids * input_mask
    inner = input_mask.bool() & (pos > 0) & (pos < (lens - 1).unsqueeze(1))
    pick = (torch.rand(B, T, generator=g) < 0.15) & inner
    pick[:, 1] = True  # at least one masked token per sample
    lm_label_ids = torch.where(pick, ids, torch.full_like(ids, -1))
    ids = torch.where(pick, torch.full_like(ids, 103 % vocab_size), ids)
    segment_ids = torch.zeros(B, T, dtype=torch.long)
    is_match = torch.randint(0, 2, (B,), generator=g)

    nbox = torch.full((B,), n_det) if full_length else torch.randint(min(10, n_det), n_det + 1, (B,), generator=g)
    feat = torch.relu(torch.randn(B, n_det, feat_dim, generator=g))
    region_pick = torch.rand(B, n_det, generator=g) < 0.15
    region_pick[:, 0] = True
    region_pick &= torch.arange(n_det).unsqueeze(0) < nbox.unsqueeze(1)
    feat = feat * (~region_pick).unsqueeze(-1).float()
    x1 = torch.rand(B, n_det, generator=g) * 0.7
    y1 = torch.rand(B, n_det, generator=g) * 0.7
    w = 0.05 + 0.25 * torch.rand(B, n_det, generator=g)
    h = 0.05 + 0.25 * torch.rand(B, n_det, generator=g)
    loc = torch.stack([x1, y1, x1 + w, y1 + h, w * h], dim=-1)[..., :num_locs]
    image_label = torch.where(region_pick, torch.ones(B, n_det, dtype=torch.long),
                              torch.full((B, n_det), -1, dtype=torch.long))
    image_cls = torch.softmax(3.0 * torch.randn(B, n_det, num_classes, generator=g), dim=-1)
    det_mask = (torch.arange(n_det).unsqueeze(0) < nbox.unsqueeze(1)).long()

    if global_first:
        keep = (det_mask.bool() & ~region_pick).float().unsqueeze(-1)
        g_feat = (feat * keep).sum(1, keepdim=True) / keep.sum(1, keepdim=True).clamp(min=1.0)
        g_loc = torch.tensor([0.0, 0.0, 1.0, 1.0, 1.0])[:num_locs].expand(B, 1, num_locs)
        feat = torch.cat([g_feat, feat], dim=1)
        loc = torch.cat([g_loc, loc], dim=1)
        image_mask = torch.cat([torch.ones(B, 1, dtype=torch.long), det_mask], dim=1)
    else:
        image_mask = det_mask

    if global_last:  # rotate the global row to the end
        feat = torch.cat([feat[:, 1:], feat[:, :1]], dim=1)
        loc = torch.cat([loc[:, 1:], loc[:, :1]], dim=1)
        image_mask = torch.cat([image_mask[:, 1:], image_mask[:, :1]], dim=1)
    out = dict(
        input_ids=ids, image_feat=feat.contiguous(), image_loc=loc.contiguous(), segment_ids=segment_ids,
        input_mask=input_mask, image_mask=image_mask, lm_label_ids=lm_label_ids,
        image_label=image_label, image_cls=image_cls.contiguous(), is_match=is_match,
    )
    if detector_labels:
        g2 = torch.Generator().manual_seed(seed + 7919)
        out.update(obj_labels=torch.randint(0, 1600, (B, n_det), generator=g2), obj_confs=torch.rand(B, n_det, generator=g2),
                   attr_labels=torch.randint(0, 400, (B, n_det), generator=g2), attr_confs=torch.rand(B, n_det, generator=g2))
    return out


def raw_region_batch(batch, region_len, *, feat_dim=2048, seed=2468, full=True):
    """Detector output BEFORE the loader touched it (what `BertPreprocessBatch.__call__` receives per sample,
    volta/datasets/concept_cap_dataset.py:429-431, stacked over the batch): pooled features (post-ReLU, fp32), boxes
    in PIXELS (x1, y1, x2, y2), image (h, w), box count, and `sel` - which regions the masking draw of
    `random_region` (:661-690) picked: 0 = untouched, 1 = predicted and zeroed, 2 = predicted, features kept.
    Boxes are drawn around a few centres so that IoU > 0.4 overlaps (the exclusion rule of :679) really occur.
    full=False: ragged box counts (rows beyond the count are zero, as the loader pads them, :435-446)."""
    g = torch.Generator().manual_seed(seed)
    B, R = batch, region_len
    hw = torch.stack([torch.randint(240, 801, (B,), generator=g), torch.randint(320, 1025, (B,), generator=g)], 1).float()
    h, w = hw[:, :1], hw[:, 1:]
    nbox = torch.full((B,), R, dtype=torch.int32) if full else \
        torch.randint(max(1, R // 3), R + 1, (B,), generator=g).to(torch.int32)
    live = torch.arange(R).unsqueeze(0) < nbox.unsqueeze(1)
    feat = torch.relu(torch.randn(B, R, feat_dim, generator=g)) * live.unsqueeze(-1)
    n_centres = max(2, R // 4)
    cx = torch.rand(B, n_centres, generator=g)
    cy = torch.rand(B, n_centres, generator=g)
    which = torch.randint(0, n_centres, (B, R), generator=g)
    bx = (cx.gather(1, which) + 0.04 * torch.randn(B, R, generator=g)).clamp(0.05, 0.95) * w
    by = (cy.gather(1, which) + 0.04 * torch.randn(B, R, generator=g)).clamp(0.05, 0.95) * h
    bw = (0.1 + 0.3 * torch.rand(B, R, generator=g)) * w
    bh = (0.1 + 0.3 * torch.rand(B, R, generator=g)) * h
    x1 = (bx - bw / 2).clamp(min=0.0)
    y1 = (by - bh / 2).clamp(min=0.0)
    x2 = torch.minimum(bx + bw / 2, w - 1)
    y2 = torch.minimum(by + bh / 2, h - 1)
    box = torch.stack([x1, y1, x2, y2], -1) * live.unsqueeze(-1)
    u = torch.rand(B, R, generator=g)
    sel = torch.where(u < 0.15 * 0.9, 1, torch.where(u < 0.15, 2, 0)).to(torch.uint8) * live.to(torch.uint8)
    return dict(raw_feat=feat.contiguous(), raw_box=box.contiguous(), image_hw=hw, num_boxes=nbox, sel=sel)


def retrieval_batch(num_pairs, seq_len, num_regions, *, vocab_size=30522, feat_dim=2048, num_locs=5, seed=4321):
    """One caption repeated against `num_pairs` candidate images (`eval_retrieval.py:172-191`)."""
    b = pretraining_batch(num_pairs, seq_len, num_regions, vocab_size=vocab_size, feat_dim=feat_dim,
                          num_locs=num_locs, seed=seed)
    for k in ("input_ids", "segment_ids", "input_mask"):
        b[k] = b[k][:1].expand(num_pairs, -1).contiguous()
    return {k: b[k] for k in ("input_ids", "image_feat", "image_loc", "segment_ids", "input_mask", "image_mask")}


def retrieval_case(num_captions, num_images, seed=0):
    """Score / target matrices in the layout of eval_retrieval.py:164-165: caption c belongs to image
    c // (num_captions // num_images) (Flickr30k: 5 captions per image); scores are distinct fp32 values with the matching
    pair pushed up, so ranks spread over the first few dozen places."""
    g = torch.Generator().manual_seed(seed)
    C, I = num_captions, num_images
    per = max(1, C // I)
    owner = (torch.arange(C) // per).clamp(max=I - 1)
    targets = torch.zeros(C, I, dtype=torch.uint8)
    targets[torch.arange(C), owner] = 1
    scores = torch.randperm(C * I, generator=g).float().reshape(C, I) / (C * I) * 8.0 - 4.0   # distinct values
    boost = torch.rand(C, generator=g) * 6.0
    scores[torch.arange(C), owner] += boost
    return scores.contiguous(), targets


def seeded_state_dict(reference_state, seed=0, std=0.02):
    """Deterministic weights for parity tests: N(0, std) for matrices / embeddings,
    N(1, std) for LayerNorm weights, N(0, std) for biases, drawn key by key in
    sorted order.  Aliased entries (same storage in `reference_state`) get the
    same tensor.  `reference_state` only supplies names, shapes and aliasing."""
    g = torch.Generator().manual_seed(seed)
    by_ptr, out = {}, {}
    for key in sorted(reference_state):
        t = reference_state[key]
        ptr = t.data_ptr()
        if ptr in by_ptr:
            out[key] = by_ptr[ptr]
            continue
        v = torch.randn(t.shape, generator=g) * std
        if key.endswith("LayerNorm.weight") or key.endswith("layer_norm.weight") or \
                (key.endswith(".weight") and t.dim() == 1):
            v = v + 1.0
        by_ptr[ptr] = out[key] = v
    return out
