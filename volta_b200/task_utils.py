"""Task criteria and scores on the device - the part of `volta/task_utils.py` that sits right
after the `BertForVLTasks` forward in `train_task.py` / `eval_task.py`:

  * `LossMap` / `LoadLoss`             (volta/task_utils.py:24-27, 284-287)
  * `compute_score_with_logits`        (:429-434)
  * the per-task-type loss + score     (`ForwardModelsTrain` :236-282, `ForwardModelsVal` :102-137) -> `task_loss_and_score`
  * the per-"process" batch reshapes   (:148-233 / :36-94)                                         -> `expand_task_batch`
  * `ForwardModelsTrain` / `ForwardModelsVal` themselves, same signatures and returns.

The reference runs ~10 small torch kernels per criterion call and materialises a [B, C] one-hot matrix for the score;
here the criteria are `vb_bce_logits_fwd/bwd` and `vb_xent_fwd/bwd`, the arg-max + label pick of every score is
`vb_argmax_pick` (no one-hot matrix), and the losses are formed in fp32 whatever the dtype of the logits.
Datasets, tokenisers and loaders (`LoadDataset*`, :290-428) are out of scope.  No CPU path: CUDA tensors only.
"""
import torch

from . import engine, ops


class _BCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        B = logits.shape[0]
        x = logits.contiguous().view(B, -1)
        t = target.to(torch.float32).contiguous().view(B, -1)
        if t.shape != x.shape:
            raise ValueError("Target size ({}) must be the same as input size ({})".format(tuple(target.shape), tuple(logits.shape)))
        row = torch.empty(B, dtype=torch.float32, device=x.device)
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        ctx.scale = 1.0 / float(x.numel())
        ops.bce_logits_fwd(x, t, row, ctx.scale, out)
        engine._count(2)
        ctx.save_for_backward(x, t)
        ctx.shape = logits.shape
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        x, t = ctx.saved_tensors
        dx = torch.empty_like(x)
        ops.bce_logits_bwd(x, t, g.reshape(1).to(torch.float32).contiguous(), ctx.scale, dx)
        engine._count()
        return dx.view(ctx.shape), None


class _XentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        if logits.dim() != 2 or target.dim() != 1 or target.shape[0] != logits.shape[0]:
            raise ValueError("CrossEntropyLoss expects logits [N, C] and class indices [N], got {} and {}".format(
                tuple(logits.shape), tuple(target.shape)))
        x = logits.contiguous()
        lab = target.to(torch.int64).contiguous()
        N, C = x.shape
        lse = torch.empty(N, dtype=torch.float32, device=x.device)
        sums = torch.zeros(2, dtype=torch.float32, device=x.device)      # [loss sum, counted rows]
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        ops.xent_fwd(x, lab, C, lse, sums[0:1], sums[1:2])
        ops.loss_finalize(sums[0:1], sums[1:2], out)
        engine._count(2)
        ctx.save_for_backward(x, lab, lse, sums)
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        x, lab, lse, sums = ctx.saved_tensors
        dx = torch.empty_like(x)
        ops.xent_bwd(x, lab, x.shape[1], lse, sums[1:2], g.reshape(1).to(torch.float32).contiguous(), dx)
        engine._count()
        return dx, None


class BCEWithLogitsLoss(torch.nn.Module):
    """nn.BCEWithLogitsLoss(reduction="mean") (volta/task_utils.py:25): mean over every element, fp32 scalar."""

    def __init__(self, reduction="mean"):
        super().__init__()
        if reduction != "mean":
            raise ValueError("volta_b200.task_utils.BCEWithLogitsLoss: only reduction='mean' (the reference's LossMap entry)")

    def forward(self, input, target):
        ops.require_cuda(input, target)
        return _BCEFn.apply(input, target)


class CrossEntropyLoss(torch.nn.Module):
    """nn.CrossEntropyLoss() (volta/task_utils.py:26): mean over the rows whose label is not ignored (negative labels -
    torch's ignore_index = -100 included - are ignored)."""

    def forward(self, input, target):
        ops.require_cuda(input, target)
        return _XentFn.apply(input, target)


LossMap = {
    "BCEWithLogitLoss": BCEWithLogitsLoss(reduction="mean"),
    "CrossEntropyLoss": CrossEntropyLoss(),
}


def LoadLoss(task_cfg, task_id):
    """volta/task_utils.py:284-287."""
    return LossMap[task_cfg["TASK" + task_id]["loss"]]


def _argmax_pick(logits, labels=None, want_arg=True, want_pick=True):
    B = logits.shape[0]
    x = logits.detach().contiguous().view(B, -1)
    lab = None if labels is None else labels.detach().to(torch.float32).contiguous().view(B, -1)
    arg = torch.empty(B, dtype=torch.int64, device=x.device) if want_arg else None
    pick = torch.empty(B, dtype=torch.float32, device=x.device) if want_pick else None
    ops.argmax_pick(x, lab, arg, pick)
    engine._count()
    return arg, pick


def compute_score_with_logits(logits, labels):
    """volta/task_utils.py:429-434: the label value at each row's arg-max prediction.  The reference returns the
    [B, C] matrix `one_hot(argmax) * labels`, of which every caller takes `.sum()`; this returns its only non-zero
    entries as [B, 1] (same sum, same per-row values), without building the one-hot matrix."""
    ops.require_cuda(logits, labels)
    _, pick = _argmax_pick(logits, labels, want_arg=False)
    return pick.view(-1, 1)


def task_loss_and_score(task_type, vil_prediction, target, criterion, batch_size, num_options=None, multi_choice_ids=None):
    """The if / elif chain over `task_cfg[task_id]["type"]` (ForwardModelsTrain :236-280, ForwardModelsVal :102-135).
    Returns (loss tensor, batch_score as a 0-dim device tensor = the number / sum of correct predictions, NOT yet divided
    by the batch size)."""
    if task_type in ("VL-classifier", "VL-classifier-GQA"):
        loss = criterion(vil_prediction, target)
        loss = loss.mean() * target.size(1)
        score = compute_score_with_logits(vil_prediction, target).sum()
    elif task_type == "VL-logit":
        vil_logit = vil_prediction.view(batch_size, num_options)
        loss = criterion(vil_logit, target)
        preds, _ = _argmax_pick(vil_logit, want_pick=False)
        score = (preds == target).sum()
    elif task_type == "V-logit":
        loss = criterion(vil_prediction, target)
        loss = loss.mean() * target.size(1)
        # select_idx = argmax over the regions; score = its target > 0.5 (:120-122)
        _, sel = _argmax_pick(vil_prediction.squeeze(2), target.squeeze(2), want_arg=False)
        score = (sel > 0.5).sum()
    elif task_type == "V-logit-mc":
        vision_logit = vil_prediction[:, 101:]  # the reference's own offset (:124, "FIXME from ViLBERT")
        vision_logit = vision_logit.squeeze(2).gather(1, multi_choice_ids)
        vision_logit = vision_logit.unsqueeze(2)
        loss = criterion(vision_logit, target)
        loss = loss.mean() * target.size(1)
        preds, _ = _argmax_pick(vision_logit.squeeze(2), want_pick=False)
        tgt, _ = _argmax_pick(target.squeeze(2), want_pick=False)
        score = (preds == tgt).sum()
    elif task_type in ("VL-binary-classifier", "VL-tri-classifier"):
        loss = criterion(vil_prediction, target)
        loss = loss.mean()
        score = compute_score_with_logits(vil_prediction, target).sum()
    else:
        raise KeyError("unknown task type %r" % (task_type,))
    return loss, score


def expand_task_batch(process, config, features, spatials, image_mask, question, input_mask, segment_ids, target=None):
    """The reshapes by `task_cfg[task_id]["process"]` (ForwardModelsTrain :148-233, ForwardModelsVal :36-94): views and
    expands only.  Returns (features, spatials, image_mask, question, input_mask, segment_ids, target, batch_size, num_options)."""
    batch_size = features.size(0)
    num_options = None
    if process in ("dialog",):
        max_num_bbox = features.size(1)
        nround = question.size(1)
        num_options = question.size(2)
        rbatch_size = batch_size * nround
        question = question.view(rbatch_size, question.size(2), question.size(3))
        target = target.view(-1)
        input_mask = input_mask.view(rbatch_size, input_mask.size(2), input_mask.size(3))
        segment_ids = segment_ids.view(rbatch_size, segment_ids.size(2), segment_ids.size(3))
        features = features.unsqueeze(1).unsqueeze(1).expand(batch_size, nround, num_options, max_num_bbox, config.v_feature_size) \
            .contiguous().view(-1, max_num_bbox, config.v_feature_size)
        spatials = spatials.unsqueeze(1).unsqueeze(1).expand(batch_size, nround, num_options, max_num_bbox, config.num_locs) \
            .contiguous().view(-1, max_num_bbox, config.num_locs)
        image_mask = image_mask.unsqueeze(1).expand(batch_size, nround, num_options, max_num_bbox).contiguous().view(-1, max_num_bbox)
        question = question.view(-1, question.size(2))
        input_mask = input_mask.view(-1, input_mask.size(2))
        segment_ids = segment_ids.view(-1, segment_ids.size(2))
        batch_size = rbatch_size
    elif process in ("expand",):
        max_num_bbox = features.size(1)
        num_options = question.size(1)
        features = features.unsqueeze(1).expand(batch_size, num_options, max_num_bbox, config.v_feature_size) \
            .contiguous().view(-1, max_num_bbox, config.v_feature_size)
        spatials = spatials.unsqueeze(1).expand(batch_size, num_options, max_num_bbox, config.num_locs) \
            .contiguous().view(-1, max_num_bbox, config.num_locs)
        image_mask = image_mask.unsqueeze(1).expand(batch_size, num_options, max_num_bbox).contiguous().view(-1, max_num_bbox)
        question = question.view(-1, question.size(2))
        input_mask = input_mask.view(-1, input_mask.size(2))
        segment_ids = segment_ids.view(-1, segment_ids.size(2))
    elif process in ("retrieval",):
        num_options = question.size(1)
        features = features.view(-1, features.size(2), features.size(3))
        spatials = spatials.view(-1, spatials.size(2), spatials.size(3))
        image_mask = image_mask.view(-1, image_mask.size(2))
        question = question.view(-1, question.size(2))
        input_mask = input_mask.view(-1, input_mask.size(2))
        segment_ids = segment_ids.view(-1, segment_ids.size(2))
    elif process in ("nlvr",):
        num_options = question.size(1)
        features = features.view(batch_size * 2, int(features.size(1) / 2), features.size(2))
        spatials = spatials.view(batch_size * 2, int(spatials.size(1) / 2), spatials.size(2))
        image_mask = image_mask.view(batch_size * 2, int(image_mask.size(1) / 2))
        question = question.repeat(1, 2).view(batch_size * 2, int(question.size(1)))
        input_mask = input_mask.repeat(1, 2).view(batch_size * 2, int(input_mask.size(1)))
        segment_ids = segment_ids.repeat(1, 2).view(batch_size * 2, int(segment_ids.size(1)))
    return features, spatials, image_mask, question, input_mask, segment_ids, target, batch_size, num_options


def _forward(config, task_cfg, device, task_id, batch, model, criterion, train):
    batch = tuple(t.cuda(device=device, non_blocking=True) for t in batch)
    multi_choice_ids = None
    if task_cfg[task_id]["type"] == "V-logit-mc":
        features, spatials, image_mask, question, target, input_mask, segment_ids, multi_choice_ids, question_id = batch
    else:
        features, spatials, image_mask, question, target, input_mask, segment_ids, question_id = batch
    process = task_cfg[task_id]["process"]
    if process in ("dialog",) and not train:
        raise NotImplementedError("dialog process for validation")
    features, spatials, image_mask, question, input_mask, segment_ids, target, batch_size, num_options = expand_task_batch(
        process, config, features, spatials, image_mask, question, input_mask, segment_ids, target)
    vil_prediction, _, _, _ = model(question, features, spatials, task_id, segment_ids, input_mask, image_mask)
    loss, score = task_loss_and_score(task_cfg[task_id]["type"], vil_prediction, target, criterion, batch_size, num_options,
                                      multi_choice_ids)
    return loss, score, batch_size


def ForwardModelsTrain(config, task_cfg, device, task_id, batch, model, criterion):
    """volta/task_utils.py:140-282: returns (loss tensor, batch_score / batch_size as a Python float - one host read,
    as in the reference's `float(...)`)."""
    loss, score, batch_size = _forward(config, task_cfg, device, task_id, batch, model, criterion, True)
    return loss, float(score) / float(batch_size)


def ForwardModelsVal(config, task_cfg, device, task_id, batch, model, criterion):
    """volta/task_utils.py:30-137: returns (float(loss), float(batch_score), batch_size)."""
    with torch.no_grad():
        loss, score, batch_size = _forward(config, task_cfg, device, task_id, batch, model, criterion, False)
    return float(loss), float(score), batch_size
