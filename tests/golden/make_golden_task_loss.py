"""Golden task losses / scores / batch reshapes, produced by the reference's OWN statements (build container only).

    python tests/golden/make_golden_task_loss.py      # writes tests/golden/task_loss.pt

`volta/task_utils.py` cannot be imported here (pytorch_transformers tokenisers at module top, `.cuda()` in every
function), so the statements are read from the file by line number, dedented and executed unchanged on seeded CPU
tensors: `LossMap` :24-27, `compute_score_with_logits` :429-434 (its `.cuda()` dropped), the per-type loss / score chain
of `ForwardModelsTrain` :237-279 and of `ForwardModelsVal` :94-135, and the per-process reshapes :149-233.
Inputs are regenerated from seeds by `task_case` / `process_case` (shared with tests/test_task_utils.py); stored are the
reference's losses, scores, gradients of the loss w.r.t. the prediction, and the reshaped batch tensors.
"""
import os
import sys
import textwrap
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/volta/task_utils.py"

# (task type, loss name) as config_tasks/all_tasks.yml pairs them
TYPES = [("VL-classifier", "BCEWithLogitLoss"), ("VL-classifier-GQA", "BCEWithLogitLoss"), ("VL-logit", "CrossEntropyLoss"),
         ("V-logit", "BCEWithLogitLoss"), ("V-logit-mc", "BCEWithLogitLoss"), ("VL-binary-classifier", "BCEWithLogitLoss"),
         ("VL-tri-classifier", "BCEWithLogitLoss")]
# "dialog" is absent: the reference's own expression for it (:172-177, a [B, V] mask expanded to 4-D) raises in torch
PROCESSES = ["normal", "expand", "retrieval", "nlvr"]


def task_case(task_type, seed=0):
    """Seeded (vil_prediction, target, batch_size, num_options, multi_choice_ids) of one task type."""
    g = torch.Generator().manual_seed(100 + seed)
    r = lambda *s: torch.randn(*s, generator=g)
    if task_type in ("VL-classifier", "VL-classifier-GQA"):
        B, C = 7, 53
        target = torch.zeros(B, C)
        for b in range(B):
            idx = torch.randperm(C, generator=g)[:3]
            target[b, idx] = torch.tensor([1.0, 0.6, 0.3])[: len(idx)] if task_type == "VL-classifier" else torch.tensor([1.0, 0.0, 0.0])
        pred = r(B, C) * 2
        pred[0, target[0].argmax()] = 9.0      # at least one correct arg-max
        return pred, target, B, None, None
    if task_type == "VL-logit":
        B, O = 5, 4
        return r(B * O, 1), torch.randint(0, O, (B,), generator=g), B, O, None
    if task_type == "V-logit":
        B, V = 6, 11
        return r(B, V, 1) * 3, (torch.rand(B, V, 1, generator=g) > 0.7).float(), B, None, None
    if task_type == "V-logit-mc":
        B, V, K = 4, 101 + 12, 4
        ids = torch.stack([torch.randperm(12, generator=g)[:K] for _ in range(B)])
        target = torch.zeros(B, K, 1)
        target[torch.arange(B), torch.randint(0, K, (B,), generator=g), 0] = 1.0
        pred = r(B, V, 1) * 2
        for b in (0, 2):                          # two correct choices
            pred[b, 101 + ids[b, int(target[b, :, 0].argmax())], 0] = 9.0
        return pred, target, B, None, ids
    C = 2 if task_type == "VL-binary-classifier" else 3
    B = 9
    target = torch.zeros(B, C)
    target[torch.arange(B), torch.randint(0, C, (B,), generator=g)] = 1.0
    return r(B, C) * 2, target, B, None, None


def process_case(process, seed=0):
    """Seeded raw batch tensors of one `process`; (features, spatials, image_mask, question, input_mask, segment_ids, target)."""
    g = torch.Generator().manual_seed(200 + seed)
    B, V, F, L, T = 3, 6, 16, 5, 7
    ri = lambda hi, *s: torch.randint(0, hi, s, generator=g)
    if process == "dialog":
        R, O = 2, 4
        return (torch.randn(B, V, F, generator=g), torch.rand(B, V, L, generator=g), ri(2, B, V), ri(50, B, R, O, T), ri(2, B, R, O, T),
                ri(2, B, R, O, T), ri(O, B, R))
    if process == "expand":
        O = 4
        return (torch.randn(B, V, F, generator=g), torch.rand(B, V, L, generator=g), ri(2, B, V), ri(50, B, O, T), ri(2, B, O, T),
                ri(2, B, O, T), ri(O, B))
    if process == "retrieval":
        O = 4
        return (torch.randn(B, O, V, F, generator=g), torch.rand(B, O, V, L, generator=g), ri(2, B, O, V), ri(50, B, O, T),
                ri(2, B, O, T), ri(2, B, O, T), ri(O, B))
    if process == "nlvr":
        return (torch.randn(B, 2 * V, F, generator=g), torch.rand(B, 2 * V, L, generator=g), ri(2, B, 2 * V), ri(50, B, T), ri(2, B, T),
                ri(2, B, T), torch.eye(2)[ri(2, B)])
    return (torch.randn(B, V, F, generator=g), torch.rand(B, V, L, generator=g), ri(2, B, V), ri(50, B, T), ri(2, B, T), ri(2, B, T),
            torch.rand(B, 9, generator=g))


def ref_lines(lo, hi):
    src = open(REF).read().split("\n")
    return textwrap.dedent("\n".join(src[lo - 1:hi]))


def main():
    import torch.nn as nn
    loss_map_src = ref_lines(24, 27)
    score_src = ref_lines(429, 434).replace(".cuda()", "")
    train_chain, val_chain = ref_lines(237, 279), ref_lines(94, 135)
    train_process = ref_lines(149, 233)
    assert loss_map_src.startswith("LossMap = {") and score_src.startswith("def compute_score_with_logits")
    assert train_chain.lstrip().startswith("# for different task") and val_chain.startswith('if task_cfg[task_id]["type"] == "VL-classifier"')
    assert train_process.startswith('if task_cfg[task_id]["process"] in ["dialog"]') and "nlvr" in train_process
    base = dict(torch=torch, nn=nn)
    exec(loss_map_src, base)
    exec(score_src, base)
    out = {"types": {}, "process": {}}
    for task_type, loss_name in TYPES:
        pred, target, B, O, ids = task_case(task_type)
        rec = {}
        for mode, chain in (("train", train_chain), ("val", val_chain)):
            p = pred.clone().requires_grad_(True)
            ns = dict(base, task_cfg={"T": {"type": task_type}}, task_id="T", criterion=base["LossMap"][loss_name], vil_prediction=p,
                      target=target.clone(), batch_size=B, num_options=O, multi_choice_ids=ids)
            exec(chain, ns)
            loss = ns["loss"]
            (grad,) = torch.autograd.grad(loss, p)
            rec[mode] = dict(loss=float(loss), score=float(ns["batch_score"]), grad=grad)
        out["types"][task_type] = dict(loss_name=loss_name, **rec)
        print(task_type, {m: (round(rec[m]["loss"], 5), rec[m]["score"]) for m in rec})
    cfg = types.SimpleNamespace(v_feature_size=16, num_locs=5)
    names = ("features", "spatials", "image_mask", "question", "input_mask", "segment_ids", "target")
    for process in PROCESSES:
        raw = process_case(process)
        ns = dict(base, task_cfg={"T": {"process": process}}, task_id="T", config=cfg, batch_size=raw[0].size(0), **dict(zip(names, raw)))
        exec(train_process, ns)
        out["process"][process] = dict({k: ns[k].clone() for k in names}, batch_size=int(ns["batch_size"]),
                                       num_options=ns.get("num_options"))
        print(process, {k: tuple(ns[k].shape) for k in names}, ns["batch_size"], ns.get("num_options"))
    path = os.path.join(HERE, "task_loss.pt")
    torch.save(out, path)
    print("->", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
