"""Region batch assembly: the numpy oracle against the outputs of the reference's own loader code
(tests/golden/pipeline_*.pt, made by tests/golden/make_golden_pipeline.py), BIT FOR BIT; the CUDA kernel against the
oracle (`-m gpu`) at the golden sizes, at ragged box counts and at the Conceptual-Captions batch shape."""
import os

import numpy as np
import pytest
import torch

from oracle.pipeline_oracle import assemble_region_batch
from volta_b200.synthetic import raw_region_batch

GOLDEN = ["pipeline_first_r8", "pipeline_last_r8", "pipeline_first_r36", "pipeline_locs4_r12", "pipeline_nomask_r10"]


def _load(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, name + ".pt"))
    B, R, add_global, num_locs, seed = g["case"]
    raw = raw_region_batch(B, R, seed=seed, full=True)
    return g, raw, add_global, num_locs


def _oracle(raw, sel, num_locs, add_global):
    return assemble_region_batch(raw["raw_feat"].numpy(), raw["raw_box"].numpy(), raw["image_hw"].numpy(),
                                 raw["num_boxes"].numpy(), sel.numpy(), num_locs, add_global)


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_equals_the_reference_loader_bit_for_bit(name, golden_dir):
    g, raw, add_global, num_locs = _load(golden_dir, name)
    feat, loc, mask, label, _ = _oracle(raw, g["sel"], num_locs, add_global)
    assert np.array_equal(feat, g["image_feat"].numpy())
    assert np.array_equal(loc, g["image_loc"].numpy())
    assert np.array_equal(mask, g["image_mask"].numpy())
    assert np.array_equal(label, g["image_label"].numpy())
    assert g["image_feat"].dtype == torch.float32 and g["image_loc"].dtype == torch.float32


def test_overlap_exclusion_changes_the_divisor():
    """The case must exercise :679 - regions overlapping a masked one leave the count."""
    raw = raw_region_batch(4, 36, seed=99, feat_dim=8)
    *_, count = _oracle(raw, raw["sel"], 5, "first")
    picked = (raw["sel"] != 0).sum(1).numpy()
    assert (36 - count >= picked).all() and (36 - count > picked).any()


def test_host_class_refuses_to_run_without_a_device():
    from volta_b200.pipeline import RegionBatchAssembler
    asm = RegionBatchAssembler(8, feature_size=16)
    raw = raw_region_batch(2, 8, feat_dim=16)
    with pytest.raises(RuntimeError):
        asm(raw["raw_feat"], raw["raw_box"], raw["image_hw"], raw["num_boxes"], raw["sel"])
    with pytest.raises(ValueError):
        RegionBatchAssembler(8, add_global_imgfeat="middle")


# ----------------------------------------------------------------------------------------------- GPU
def _device_run(raw, sel, num_locs, add_global, out_dtype, in_dtype=torch.float32):
    from volta_b200.pipeline import RegionBatchAssembler
    B, R, F = raw["raw_feat"].shape
    asm = RegionBatchAssembler(R, feature_size=F, num_locs=num_locs, add_global_imgfeat=add_global, out_dtype=out_dtype)
    dev = torch.device("cuda:0")
    out = asm(raw["raw_feat"].to(in_dtype).pin_memory(), raw["raw_box"].pin_memory(), raw["image_hw"].pin_memory(),
              raw["num_boxes"].pin_memory(), sel.pin_memory(), device=dev)
    torch.cuda.synchronize()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN)
def test_kernel_equals_the_reference_loader_bit_for_bit(name, golden_dir):
    g, raw, add_global, num_locs = _load(golden_dir, name)
    feat, loc, mask, label, _ = _device_run(raw, g["sel"], num_locs, add_global, torch.float32)
    assert torch.equal(feat.cpu(), g["image_feat"])
    assert torch.equal(loc.cpu(), g["image_loc"])
    assert torch.equal(mask.cpu(), g["image_mask"])
    assert torch.equal(label.cpu(), g["image_label"])


@pytest.mark.gpu
@pytest.mark.parametrize("B,R,F,full,add_global,num_locs", [
    (5, 36, 2048, False, "first", 5), (3, 100, 2048, False, "last", 5), (2, 7, 40, False, "first", 4),
    (1, 1, 8, True, "first", 5), (3, 9, 30, False, "first", 5), (4, 12, 1024, False, "none", 5), (256, 36, 2048, True, "first", 5)])
def test_kernel_equals_the_oracle(B, R, F, full, add_global, num_locs):
    raw = raw_region_batch(B, R, feat_dim=F, seed=100 + B + R, full=full)
    feat, loc, mask, label, count = _device_run(raw, raw["sel"], num_locs, add_global, torch.float32)
    o_feat, o_loc, o_mask, o_label, o_count = _oracle(raw, raw["sel"], num_locs, add_global)
    assert np.array_equal(feat.cpu().numpy(), o_feat)
    assert np.array_equal(loc.cpu().numpy(), o_loc)
    assert np.array_equal(mask.cpu().numpy(), o_mask)
    assert np.array_equal(label.cpu().numpy(), o_label)
    assert np.array_equal(count.cpu().numpy().astype(np.int64), o_count)


@pytest.mark.gpu
def test_empty_batch_returns_empty_outputs():
    raw = raw_region_batch(2, 8, feat_dim=16)
    from volta_b200.pipeline import RegionBatchAssembler
    raw = {k: v[:0].cuda() for k, v in raw.items()}
    asm = RegionBatchAssembler(8, feature_size=16, out_dtype=torch.float32)
    feat, loc, mask, label, count = asm(raw["raw_feat"], raw["raw_box"], raw["image_hw"], raw["num_boxes"], raw["sel"])
    assert feat.shape == (0, 9, 16) and loc.shape == (0, 9, 5) and mask.shape == (0, 9) and label.shape == (0, 8) and count.numel() == 0


@pytest.mark.gpu
def test_bf16_output_is_the_rounded_fp32_result_and_bf16_upload_is_accepted():
    raw = raw_region_batch(6, 36, seed=5, full=False)
    o_feat = torch.from_numpy(_oracle(raw, raw["sel"], 5, "first")[0])
    feat = _device_run(raw, raw["sel"], 5, "first", torch.bfloat16)[0]
    assert feat.dtype == torch.bfloat16 and torch.equal(feat.cpu(), o_feat.to(torch.bfloat16))
    # bf16 upload (half the H2D bytes): the oracle sees the same rounded inputs
    raw16 = dict(raw, raw_feat=raw["raw_feat"].to(torch.bfloat16).float())
    o16 = torch.from_numpy(_oracle(raw16, raw["sel"], 5, "first")[0])
    feat16 = _device_run(raw, raw["sel"], 5, "first", torch.bfloat16, in_dtype=torch.bfloat16)[0]
    assert torch.equal(feat16.cpu(), o16.to(torch.bfloat16))


@pytest.mark.gpu
def test_assembled_batch_feeds_the_model_like_the_host_assembled_one():
    """image_feat / image_loc / image_mask from the device assembler drive BertForVLPreTraining to the same losses as
    the same batch assembled by the oracle on the host."""
    from volta_b200.config import BertConfig, named_config_dict
    from volta_b200.encoders import BertForVLPreTraining
    from volta_b200.synthetic import pretraining_batch
    d = named_config_dict("ctrl_vilbert_base")
    d.update(num_hidden_layers=d.get("num_hidden_layers", 12))
    torch.manual_seed(0)
    model = BertForVLPreTraining(BertConfig.from_dict(d)).cuda().eval()
    B, T, R = 4, 12, 36
    raw = raw_region_batch(B, R, seed=21, full=False)
    text = pretraining_batch(B, T, R + 1, seed=3)
    losses = []
    for src in ("device", "host"):
        if src == "device":
            feat, loc, mask, label, _ = _device_run(raw, raw["sel"], 5, "first", torch.float32)
        else:
            feat, loc, mask, label, _ = [torch.from_numpy(x).cuda() for x in _oracle(raw, raw["sel"], 5, "first")]
        with torch.no_grad():
            out = model(text["input_ids"].cuda(), feat, loc, text["segment_ids"].cuda(), text["input_mask"].cuda(), mask,
                        text["lm_label_ids"].cuda(), label, text["image_cls"].cuda(), None, None, None, None, None,
                        text["is_match"].cuda())
        losses.append([float(x) for x in out[:3]])
    for a, b in zip(*losses):   # the loss sums use fp32 atomics: last-bit run-to-run differences
        assert abs(a - b) <= 2e-6 * abs(b), losses


@pytest.mark.gpu
@pytest.mark.parametrize("pinned,wire", [(False, None), (True, torch.bfloat16)])
def test_prefetching_loader_yields_the_assembled_batches_in_order(pinned, wire):
    """PrefetchingRegionLoader (two slots, staging thread, copy stream): every yielded batch equals the direct assembly
    of the same host batch, in order, while the consumer still holds the previous one; device buffers are reused (a slot's
    addresses repeat every `depth` batches); pass-through tensors arrive unchanged."""
    from volta_b200.pipeline import PrefetchingRegionLoader, RegionBatchAssembler
    from volta_b200.synthetic import pretraining_batch
    B, R, F, n = 16, 36, 256, 7
    dev = torch.device("cuda:0")
    host = []
    for i in range(n):
        raw = raw_region_batch(B, R, feat_dim=F, seed=300 + i, full=False)
        text = pretraining_batch(B, 12, R + 1, feat_dim=F, seed=400 + i)
        hb = {k: raw[k] for k in ("raw_feat", "raw_box", "image_hw", "num_boxes", "sel")}
        hb.update(input_ids=text["input_ids"], image_cls=text["image_cls"], is_match=text["is_match"])
        host.append({k: (v.pin_memory() if pinned else v) for k, v in hb.items()})
    asm = RegionBatchAssembler(R, feature_size=F, num_locs=5, add_global_imgfeat="first", out_dtype=torch.bfloat16)
    loader = PrefetchingRegionLoader(iter(host), asm, dev, depth=2, wire_dtype=wire)
    # The consumer never synchronises inside the loop: it queues a delay and then a read of the batch on its stream, and
    # asks for the next batch right away.  A slot may only be overwritten after that queued read has run (the `left`
    # event), and a batch may only be read after its upload + assembly have finished (the `ready` event).
    ptrs, reads = [], []
    for i, batch in enumerate(loader):
        torch.cuda._sleep(5_000_000)
        reads.append({k: batch[k].clone() for k in ("image_feat", "image_loc", "image_mask", "image_label", "num_regions",
                                                     "input_ids", "image_cls", "is_match")})
        ptrs.append(batch["image_feat"].data_ptr())
    torch.cuda.synchronize()
    assert len(reads) == n
    for i, got in enumerate(reads):
        hb = host[i]
        rf = hb["raw_feat"] if wire is None else hb["raw_feat"].to(wire)
        feat, loc, mask, label, count = asm(rf.cuda(), hb["raw_box"].cuda(), hb["image_hw"].cuda(), hb["num_boxes"].cuda(), hb["sel"].cuda())
        torch.cuda.synchronize()
        assert torch.equal(got["image_feat"], feat) and torch.equal(got["image_loc"], loc), i
        assert torch.equal(got["image_mask"], mask) and torch.equal(got["image_label"], label), i
        assert torch.equal(got["num_regions"], count), i
        for k in ("input_ids", "image_cls", "is_match"):
            assert torch.equal(got[k].cpu(), hb[k]), (i, k)
    assert len(set(ptrs)) == 2 and ptrs[0] == ptrs[2] == ptrs[4] and ptrs[1] == ptrs[3]
    assert loader.h2d_bytes > 0
