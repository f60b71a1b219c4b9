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
