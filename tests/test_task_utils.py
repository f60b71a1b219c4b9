"""volta_b200.task_utils (task criteria, scores, batch reshapes) against the reference's OWN
statements executed on seeded tensors (tests/golden/make_golden_task_loss.py: volta/task_utils.py:24-27, 94-135, 149-233,
237-279, 429-434).  CPU: the module's logic with the three kernels emulated from their documented semantics; GPU: the
kernels, fp32 (1e-5) and bf16 logits (loss 2e-2: the logits are rounded before the criterion sees them)."""
import os
import types

import pytest
import torch

import emu_ops
from golden_util import ROOT  # (puts tests/golden on sys.path)
import make_golden_task_loss as G  # noqa: E402  (seeded cases; does not import the reference)
from volta_b200 import task_utils as TU


def _gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "task_loss.pt"), weights_only=False)


def _check_types(device, dtype, tol):
    gold = _gold()
    for task_type, loss_name in G.TYPES:
        ref = gold["types"][task_type]
        assert ref["loss_name"] == loss_name
        pred, target, B, O, ids = G.task_case(task_type)
        pred = pred.to(device=device, dtype=dtype).requires_grad_(True)
        target = target.to(device)
        ids = None if ids is None else ids.to(device)
        criterion = TU.LossMap[loss_name]
        loss, score = TU.task_loss_and_score(task_type, pred, target, criterion, B, O, ids)
        assert loss.dim() == 0 and loss.dtype == torch.float32
        assert abs(float(loss) - ref["train"]["loss"]) <= tol * max(1.0, abs(ref["train"]["loss"])), (task_type, float(loss), ref["train"]["loss"])
        if dtype == torch.float32:       # arg-max decisions are exact in fp32; bf16 rounding may flip near-ties
            assert float(score) == ref["val"]["score"], (task_type, float(score), ref["val"]["score"])
            assert abs(float(score) / B - ref["train"]["score"]) < 1e-6
        (grad,) = torch.autograd.grad(loss, pred)
        g = ref["train"]["grad"]
        err = float((grad.float().cpu() - g).norm() / g.norm())
        assert err < (1e-5 if dtype == torch.float32 else 2e-2), (task_type, err)


def test_task_losses_and_scores_match_reference_statements_emulated(monkeypatch):
    emu_ops.install(monkeypatch)
    _check_types("cpu", torch.float32, 1e-6)


@pytest.mark.parametrize("process", G.PROCESSES)
def test_task_batch_reshapes_match_reference_statements(process):
    gold = _gold()["process"][process]
    raw = G.process_case(process)
    cfg = types.SimpleNamespace(v_feature_size=16, num_locs=5)
    f, s, m, q, im, sg, t, bs, no = TU.expand_task_batch(process, cfg, raw[0], raw[1], raw[2], raw[3], raw[4], raw[5], raw[6])
    for name, got in zip(("features", "spatials", "image_mask", "question", "input_mask", "segment_ids", "target"), (f, s, m, q, im, sg, t)):
        assert torch.equal(got, gold[name]), (process, name)
    assert bs == gold["batch_size"] and no == gold["num_options"]


def test_dialog_reshape_fails_like_the_reference():
    """volta/task_utils.py:172-177 expands a [B, V] mask to four dimensions, which torch rejects; same expression here."""
    g = torch.Generator().manual_seed(0)
    B, V, F, L, T, R, O = 3, 6, 16, 5, 7, 2, 4
    cfg = types.SimpleNamespace(v_feature_size=F, num_locs=L)
    with pytest.raises(RuntimeError):
        TU.expand_task_batch("dialog", cfg, torch.randn(B, V, F, generator=g), torch.rand(B, V, L, generator=g), torch.ones(B, V),
                             torch.zeros(B, R, O, T, dtype=torch.long), torch.ones(B, R, O, T, dtype=torch.long),
                             torch.zeros(B, R, O, T, dtype=torch.long), torch.zeros(B, R, dtype=torch.long))


def test_load_loss_and_errors(monkeypatch):
    emu_ops.install(monkeypatch)
    cfg = {"TASK1": {"loss": "BCEWithLogitLoss"}, "TASK8": {"loss": "CrossEntropyLoss"}}
    assert TU.LoadLoss(cfg, "1") is TU.LossMap["BCEWithLogitLoss"] and TU.LoadLoss(cfg, "8") is TU.LossMap["CrossEntropyLoss"]
    with pytest.raises(ValueError):
        TU.LossMap["BCEWithLogitLoss"](torch.zeros(2, 3), torch.zeros(2, 4))
    with pytest.raises(KeyError):
        TU.task_loss_and_score("nope", torch.zeros(2, 3), torch.zeros(2, 3), TU.LossMap["BCEWithLogitLoss"], 2)
    # an ignored label (negative) does not count in the cross-entropy mean, as with nn.CrossEntropyLoss(ignore_index=-100)
    x = torch.randn(4, 5)
    y = torch.tensor([1, -100, 3, 0])
    got = TU.LossMap["CrossEntropyLoss"](x, y)
    assert abs(float(got) - float(torch.nn.CrossEntropyLoss()(x, y))) < 1e-6


def test_no_cpu_path():
    with pytest.raises(RuntimeError, match="no CPU path"):
        TU.LossMap["BCEWithLogitLoss"](torch.zeros(2, 3), torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        TU.compute_score_with_logits(torch.zeros(2, 3), torch.zeros(2, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_task_losses_and_scores_match_reference_statements_gpu(dtype, tol):
    _check_types("cuda", dtype, tol)


@pytest.mark.gpu
def test_task_kernels_at_vqa_size_gpu():
    """B = 256 x 3129 answers (VQA): loss / gradient / score against torch on the same device tensors; ties pick the first index."""
    g = torch.Generator().manual_seed(5)
    B, C = 256, 3129
    x = (torch.randn(B, C, generator=g) * 3).cuda().requires_grad_(True)
    t = (torch.rand(B, C, generator=g) > 0.999).float().cuda() * 0.6
    loss = TU.LossMap["BCEWithLogitLoss"](x, t)
    ref = torch.nn.BCEWithLogitsLoss(reduction="mean")(x, t)
    assert abs(float(loss) - float(ref)) < 1e-6 * max(1.0, float(ref))
    (g1,) = torch.autograd.grad(loss * 7.0, x)
    (g2,) = torch.autograd.grad(ref * 7.0, x)
    assert float((g1 - g2).abs().max()) < 1e-9 + 1e-5 * float(g2.abs().max())
    sc = TU.compute_score_with_logits(x, t)
    one_hot = torch.zeros_like(t).scatter_(1, x.argmax(1, keepdim=True), 1)
    assert sc.shape == (B, 1) and float(sc.sum()) == float((one_hot * t).sum())
    y = torch.zeros(3, 9).cuda()
    y[:, [2, 5]] = 4.0
    arg, _ = TU._argmax_pick(y, want_pick=False)
    assert arg.tolist() == [2, 2, 2]
    # two launches only for the loss, one for its backward, one per score
    from volta_b200 import engine
    n0 = engine.launch_count
    TU.LossMap["BCEWithLogitLoss"](x.detach(), t)
    assert engine.launch_count - n0 == 2


# ---- ForwardModelsTrain / ForwardModelsVal around a real BertForVLTasks -------------------------------------------------
def _forward_models_case(device, precision, tol):
    """`ForwardModelsTrain` / `ForwardModelsVal` with the mini ctrl_vilbert task model: the loss equals the reference
    criterion applied to the UNMODIFIED reference's predictions (tests/golden/tasks_mini_ctrl_vilbert.pt), the score
    its arg-max count, and the loss back-propagates into encoder and head."""
    import copy
    import volta_b200
    import make_golden_tasks as MK
    from volta_b200.synthetic import seeded_state_dict
    g = torch.load(os.path.join(ROOT, "tests", "golden", "tasks_mini_ctrl_vilbert.pt"), weights_only=False)
    cfg = volta_b200.BertConfig.from_dict(g["config"])
    task_cfg = copy.deepcopy(g["task_cfg"])
    proc = {"TASK1": "normal", "TASK13": "normal", "TASK10": "normal", "TASK8": "retrieval"}
    loss_of = {"TASK1": "BCEWithLogitLoss", "TASK13": "BCEWithLogitLoss", "TASK10": "BCEWithLogitLoss", "TASK8": "CrossEntropyLoss"}
    model = volta_b200.BertForVLTasks(cfg, task_cfg, g["tasks"])
    model.load_state_dict(seeded_state_dict(model.state_dict(), seed=0))
    model.set_precision(precision).eval().to(device)
    b = MK.case_batch("tasks_mini_ctrl_vilbert")
    B = b["input_ids"].shape[0]
    gen = torch.Generator().manual_seed(3)
    for t in proc:
        task_cfg[t].update(process=proc[t], loss=loss_of[t])
        ref_pred = g["pred"][t]
        feats, locs, vmask, ids, imask, seg = b["image_feat"], b["image_loc"], b["image_mask"], b["input_ids"], b["input_mask"], b["segment_ids"]
        if t == "TASK8":       # one caption against B images: [1, B, ...] batch, class index target
            feats, locs, vmask, ids, imask, seg = (x.unsqueeze(0) for x in (feats, locs, vmask, ids, imask, seg))
            target = torch.tensor([2])
            ref_loss = torch.nn.CrossEntropyLoss()(ref_pred.view(1, B), target)
            ref_score = float((ref_pred.view(1, B).argmax(1) == target).sum())
            bs = 1
        elif t == "TASK10":
            target = (torch.rand(B, ref_pred.shape[1], 1, generator=gen) > 0.6).float()
            ref_loss = torch.nn.BCEWithLogitsLoss()(ref_pred, target) * target.size(1)
            ref_score = float((target.squeeze(2).gather(1, ref_pred.argmax(1).view(-1, 1)) > 0.5).sum())
            bs = B
        else:
            C = ref_pred.shape[1]
            target = torch.zeros(B, C)
            target[torch.arange(B), torch.randint(0, C, (B,), generator=gen)] = 1.0
            ref_loss = torch.nn.BCEWithLogitsLoss()(ref_pred, target) * (C if t == "TASK1" else 1)
            ref_score = float((torch.zeros(B, C).scatter_(1, ref_pred.argmax(1, keepdim=True), 1) * target).sum())
            bs = B
        batch = (feats, locs, vmask, ids, target, imask, seg, torch.arange(B))
        criterion = TU.LoadLoss({"TASK" + t[4:]: task_cfg[t]}, t[4:])
        model.zero_grad(set_to_none=True)
        loss, score = TU.ForwardModelsTrain(cfg, task_cfg, device, t, batch, model, criterion)
        assert abs(float(loss) - float(ref_loss)) <= tol * max(1.0, abs(float(ref_loss))), (t, float(loss), float(ref_loss))
        if precision == "fp32":
            assert abs(score - ref_score / bs) < 1e-6, (t, score, ref_score, bs)
        loss.backward()
        touched = [k for k, p in model.named_parameters() if p.grad is not None and float(p.grad.abs().sum()) > 0]
        assert any(k.startswith("bert.encoder.layer.0.") for k in touched) and any(k.startswith("clfs_dict.%s." % t) for k in touched), t
        vloss, vscore, vbs = TU.ForwardModelsVal(cfg, task_cfg, device, t, batch, model, criterion)
        assert isinstance(vloss, float) and abs(vloss - float(ref_loss)) <= tol * max(1.0, abs(float(ref_loss))) and vbs == bs
        if precision == "fp32":
            assert vscore == ref_score


def test_forward_models_train_val_emulated(monkeypatch):
    emu_ops.install(monkeypatch)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)    # the batch transfer of :31 / :141, on a CPU-only box
    _forward_models_case("cpu", "fp32", 2e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 3e-2)])
def test_forward_models_train_val_gpu(precision, tol):
    _forward_models_case("cuda", precision, tol)
