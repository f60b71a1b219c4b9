"""Golden vectors for BertForVLTasks (reference: volta/encoders.py:1117-1206) from the UNMODIFIED reference.

    python tests/golden/make_golden_tasks.py      # build container only; writes tests/golden/tasks_*.pt

Every task-head type the reference defines is covered on a mini ctrl_vilbert encoder (VL-classifier, NLVR pair
classifier, tri-classifier, retrieval VL-logit, V-logit with 1 and 2 layers); the retrieval head is also run on the
full-size vilbert_base encoder with the Flickr30k retrieval shape of BASELINE.json (T=30, V=101).  eval() mode, so
the reference's dropouts are identities.  Weights and inputs are regenerated from seeds by volta_b200.synthetic.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import MINI, import_reference, weights_checksum  # noqa: E402
from volta_b200.config import named_config_dict  # noqa: E402
from volta_b200.synthetic import pretraining_batch, seeded_state_dict  # noqa: E402

TASK_CFG = {
    "TASK1": {"type": "VL-classifier", "num_labels": 10},
    "TASK12": {"type": "VL-binary-classifier"},
    "TASK13": {"type": "VL-tri-classifier"},
    "TASK8": {"type": "VL-logit"},
    "TASK10": {"type": "V-logit"},
    "TASK11": {"type": "V-logit-2", "num_clf_layers": 2},
}
CASES = {  # name -> (base config, overrides, B, T, V, task ids)
    "tasks_mini_ctrl_vilbert": ("ctrl_vilbert_base", MINI, 4, 10, 7, list(TASK_CFG)),
    "tasks_mini_ctrl_visualbert": ("ctrl_visualbert_base", MINI, 4, 10, 7, ["TASK1", "TASK8", "TASK10"]),
    "tasks_full_vilbert_retrieval": ("vilbert_base", {}, 3, 30, 101, ["TASK8"]),
}


def case_config_dict(name):
    base, over, B, T, V, tasks = CASES[name]
    d = named_config_dict(base)
    d.update(over)
    for k in ("hidden_dropout_prob", "attention_probs_dropout_prob", "v_hidden_dropout_prob", "v_attention_probs_dropout_prob"):
        d[k] = 0.0
    return d, (B, T, V), tasks


def case_batch(name):
    d, (B, T, V), _ = case_config_dict(name)
    return pretraining_batch(B, T, V, vocab_size=d["vocab_size"], feat_dim=d["v_feature_size"], num_locs=d["num_locs"], seed=91)


def main():
    BertConfig, _ = import_reference()
    from volta.encoders import BertForVLTasks
    for name in CASES:
        d, (B, T, V), tasks = case_config_dict(name)
        torch.manual_seed(0)
        model = BertForVLTasks(BertConfig.from_dict(d), TASK_CFG, tasks)
        state = seeded_state_dict(model.state_dict(), seed=0)
        model.load_state_dict(state)
        model.eval()
        b = case_batch(name)
        out = {"config": d, "shape": (B, T, V), "tasks": tasks, "task_cfg": TASK_CFG, "weights_checksum": weights_checksum(state),
               "state_keys": sorted(state.keys()), "pred": {}}
        with torch.no_grad():
            for t in tasks:
                vil, _, _, _ = model(b["input_ids"], b["image_feat"].clone(), b["image_loc"], t, b["segment_ids"], b["input_mask"],
                                     b["image_mask"].float())
                out["pred"][t] = vil.clone()
        path = os.path.join(HERE, name + ".pt")
        torch.save(out, path)
        print("%-30s %s  %.1f KB" % (name, {t: tuple(v.shape) for t, v in out["pred"].items()}, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
