"""Golden attention data from the UNMODIFIED reference with config.visualization = True (build container only).

    python tests/golden/make_golden_attn.py       # writes tests/golden/attn_*.pt

What `BertModel.forward(..., output_all_attention_masks=True)` returns per attention sub-layer when the config asks for
visualisation (volta/encoders.py:342-356, 873-875): the (t_attn_data, v_attn_data) dicts with the intra- / inter-modal
probability blocks after the (joint) softmax, queries and keys.  Stored: the probability blocks of every attention
sub-layer (eval mode) and the queries / keys of the first and the last one.  Weights and inputs are regenerated from
seeds (tests/golden/make_golden.py).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402
from volta_b200.synthetic import seeded_state_dict  # noqa: E402

CASES = ["mini_ctrl_vilbert", "mini_ctrl_visualbert", "mini_vilbert"]


def run_case(name, BertConfig, BertForVLPreTraining):
    d, (B, T, V) = G.case_config_dict(name)
    d = dict(d, visualization=True)
    cfg = BertConfig.from_dict(d)
    model = BertForVLPreTraining(cfg)
    model.load_state_dict(seeded_state_dict(model.state_dict(), seed=0))
    model.eval()
    b = G.case_batch(name)
    with torch.no_grad():
        _, _, _, _, (at, av) = model.bert(b["input_ids"], b["image_feat"].clone(), b["image_loc"], b["segment_ids"],
                                          b["input_mask"], b["image_mask"], output_all_attention_masks=True)
    assert len(at) == len(av) and len(at) > 0

    def pick(dct, full):
        if dct is None:
            return None
        keys = ("intra_attn", "inter_attn") + (("queries", "keys") if full else ())
        return {k: (None if dct[k] is None else dct[k].clone()) for k in keys}
    n = len(at)
    return {"config": d, "shape": (B, T, V), "num_attention_sublayers": n,
            "t": [pick(x, i in (0, n - 1)) for i, x in enumerate(at)], "v": [pick(x, i in (0, n - 1)) for i, x in enumerate(av)]}


if __name__ == "__main__":
    BertConfig, BertForVLPreTraining = G.import_reference()
    for name in CASES:
        out = run_case(name, BertConfig, BertForVLPreTraining)
        path = os.path.join(HERE, "attn_" + name + ".pt")
        torch.save(out, path)
        print(name, out["num_attention_sublayers"], "attention sub-layers ->", path, os.path.getsize(path) // 1024, "KB")
