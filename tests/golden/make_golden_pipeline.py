"""Golden vectors for the region batch assembly, produced by the UNMODIFIED reference loader code (build container only).

    python tests/golden/make_golden_pipeline.py      # writes tests/golden/pipeline_*.pt

`volta/datasets/concept_cap_dataset.py` is loaded from /root/reference as a single module (its package `__init__`
pulls in every dataset and their absent dependencies); the only shims are `sys.modules` stubs for `tensorpack` and
`msgpack_numpy` (imported at the top of the file, `:13,18-19`, used only by the LMDB reader) and a toy tokenizer.
`BertPreprocessBatch.__call__` (`:429-504`) runs per sample on seeded detector outputs
(`volta_b200.synthetic.raw_region_batch`), the results are stacked the way `td.BatchData` does, and
`ConceptCapLoaderTrain.__iter__` (`:229-286`) adds the global feature.  Inputs are NOT stored (regenerated from the seed);
stored are the reference's outputs and the masking draws recovered from them (`sel`).
"""
import importlib.util
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from volta_b200.synthetic import raw_region_batch  # noqa: E402

# name -> (B, region_len, add_global_imgfeat, num_locs, data seed, python-random seed)
CASES = {
    "pipeline_first_r8": (4, 8, "first", 5, 11, 5),
    "pipeline_last_r8": (3, 8, "last", 5, 12, 6),
    "pipeline_first_r36": (2, 36, "first", 5, 13, 7),
    "pipeline_locs4_r12": (3, 12, "first", 4, 14, 8),
    # visualization=True: random_region / random_word never mask (:625, :672) - no picked region, divisor = region_len
    "pipeline_nomask_r10": (3, 10, "first", 5, 15, 9),
}


def load_reference_module():
    td = types.ModuleType("tensorpack.dataflow")
    tp = types.ModuleType("tensorpack")
    tp.dataflow = td
    mp = types.ModuleType("msgpack_numpy")
    mp.patch = lambda: None
    sys.modules.setdefault("tensorpack", tp)
    sys.modules.setdefault("tensorpack.dataflow", td)
    sys.modules.setdefault("msgpack_numpy", mp)
    path = "/root/reference/volta/datasets/concept_cap_dataset.py"
    spec = importlib.util.spec_from_file_location("ref_concept_cap_dataset", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class ToyTokenizer:
    mask_token = "[MASK]"

    def encode(self, caption):
        return [1000 + (ord(c) % 97) for c in caption][:20]

    def convert_tokens_to_ids(self, tok):
        return 103

    def __len__(self):
        return 30522

    def add_special_tokens_single_sentence(self, tokens):
        return [101] + tokens + [102]


class FakeFlow:
    def __init__(self, batch):
        self.batch = batch

    def get_data(self):
        yield self.batch

    def size(self):
        return 1


def run_reference(mod, raw, region_len, add_global, num_locs, py_seed, visualization=False):
    pp = object.__new__(mod.BertPreprocessBatch)           # the constructor only reads the caption file
    pp.split, pp.seq_len, pp.region_len = "Train", 16, region_len
    pp.tokenizer, pp.num_caps, pp.captions = ToyTokenizer(), 1, ["a caption"]
    pp.visualization, pp.objective, pp.bert_model, pp.num_locs = visualization, 2, "bert-base-uncased", num_locs
    random.seed(py_seed)
    np.random.seed(py_seed)
    B = raw["raw_feat"].shape[0]
    per_sample = []
    for b in range(B):
        nb = int(raw["num_boxes"][b])
        feat = raw["raw_feat"][b, :nb].numpy().copy()
        box = raw["raw_box"][b, :nb].numpy().copy()
        h, w = float(raw["image_hw"][b, 0]), float(raw["image_hw"][b, 1])
        data = (feat, np.zeros((nb, 1601), np.float32), np.zeros(nb, np.int64), np.zeros(nb, np.float32),
                np.zeros(nb, np.int64), np.zeros(nb, np.float32), np.zeros((nb, 401), np.float32),
                box, nb, h, w, b, "caption number %d" % b)
        per_sample.append(pp(data))
    batch = [np.stack([np.asarray(s[k]) for s in per_sample]) for k in range(len(per_sample[0]))]   # td.BatchData
    ld = object.__new__(mod.ConceptCapLoaderTrain)
    ld.ds, ld.add_global_imgfeat, ld.num_locs = FakeFlow(batch), add_global, num_locs
    out = next(iter(ld))
    (_, _, _, _, _, image_feat, image_loc, _, _, _, _, _, _, image_label, image_mask, _) = out
    return image_feat, image_loc, image_mask, image_label


def main():
    mod = load_reference_module()
    for name, (B, R, add_global, num_locs, seed, py_seed) in CASES.items():
        raw = raw_region_batch(B, R, seed=seed, full=True)   # the reference needs num_boxes == region_len (:679 broadcast)
        nomask = "nomask" in name
        feat, loc, mask, label = run_reference(mod, raw, R, add_global, num_locs, py_seed, visualization=nomask)
        body = feat[:, 1:] if add_global == "first" else feat[:, :-1]
        zeroed = (body == 0).all(-1) & (raw["raw_feat"] != 0).any(-1)
        sel = torch.where(label == 1, torch.where(zeroed, 1, 2), 0).to(torch.uint8)
        assert (int((sel != 0).sum()) > 0) != nomask, "masking does not match the case - pick another seed"
        torch.save(dict(case=(B, R, add_global, num_locs, seed), sel=sel, image_feat=feat, image_loc=loc,
                        image_mask=mask, image_label=label), os.path.join(HERE, name + ".pt"))
        print(name, tuple(feat.shape), "masked:", int((sel == 1).sum()), "kept:", int((sel == 2).sum()))


if __name__ == "__main__":
    main()
