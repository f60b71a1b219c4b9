"""Task criteria and scores on the device - the part of `volta/task_utils.py` that sits right
after the `BertForVLTasks` forward in `train_task.py` / `eval_task.py`:

  * `LossMap` / `LoadLoss`             (volta/task_utils.py:24-27, 284-287)
  * `compute_score_with_logits`        (:429-434)
  * the per-task-type loss + score     (`ForwardModelsTrain` :237-279, `ForwardModelsVal` :94-135)  -> `task_loss_and_score`
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


# task types whose criterion value is multiplied by the number of classes / regions (`loss.mean() * target.size(1)`)
_TIMES_WIDTH = ("VL-classifier", "VL-classifier-GQA", "V-logit", "V-logit-mc")
_TASK_TYPES = _TIMES_WIDTH + ("VL-logit", "VL-binary-classifier", "VL-tri-classifier")


def task_loss_and_score(task_type, vil_prediction, target, criterion, batch_size, num_options=None, multi_choice_ids=None):
    """Loss and score of one task type - the if / elif chain over `task_cfg[task_id]["type"]` (ForwardModelsTrain
    :237-279, ForwardModelsVal :94-135).  Returns (loss tensor, score): score is a 0-dim device tensor, the number (or
    label mass) of correct predictions, NOT yet divided by the batch size.
      VL-logit                      : one of `num_options` candidates per example; class-index target; hit = arg-max == target
      VL-classifier(-GQA), binary,
      tri-classifier                : soft / one-hot label matrix; score = label at the arg-max class
      V-logit                       : one logit per region; score = the arg-max region's target > 0.5
      V-logit-mc                    : the logits of the listed regions (after the reference's offset 101, :124); hit = arg-max
                                      of the prediction == arg-max of the target"""
    if task_type not in _TASK_TYPES:
        raise KeyError("unknown task type %r" % (task_type,))
    if task_type == "VL-logit":
        logits = vil_prediction.view(batch_size, num_options)
        return criterion(logits, target), (_argmax_pick(logits, want_pick=False)[0] == target).sum()
    logits = vil_prediction
    if task_type == "V-logit-mc":
        logits = vil_prediction[:, 101:].squeeze(2).gather(1, multi_choice_ids).unsqueeze(2)
    loss = criterion(logits, target).mean()
    if task_type in _TIMES_WIDTH:
        loss = loss * target.size(1)
    if task_type == "V-logit":
        score = (_argmax_pick(logits.squeeze(2), target.squeeze(2), want_arg=False)[1] > 0.5).sum()
    elif task_type == "V-logit-mc":
        score = (_argmax_pick(logits.squeeze(2), want_pick=False)[0] == _argmax_pick(target.squeeze(2), want_pick=False)[0]).sum()
    else:
        score = compute_score_with_logits(logits, target).sum()
    return loss, score


def _per_option(x, n):
    """[B, ...] -> [B * n, ...]: every sample repeated once per answer option (rows b*n .. b*n + n-1 are sample b)."""
    return x.unsqueeze(1).expand(x.shape[0], n, *x.shape[1:]).reshape(x.shape[0] * n, *x.shape[1:])


def _merge_options(x):
    """[B, O, ...] -> [B * O, ...]."""
    return x.reshape(x.shape[0] * x.shape[1], *x.shape[2:])


def _split_pair(x):
    """[B, 2 * V, ...] -> [2 * B, V, ...]: the two images of an NLVR2 example become consecutive samples."""
    return x.reshape(x.shape[0] * 2, x.shape[1] // 2, *x.shape[2:])


def expand_task_batch(process, config, features, spatials, image_mask, question, input_mask, segment_ids, target=None):
    """The batch layouts by `task_cfg[task_id]["process"]` (ForwardModelsTrain :148-233, ForwardModelsVal :36-94), as
    views / expands of the loader's tensors.  Returns (features, spatials, image_mask, question, input_mask, segment_ids,
    target, batch_size, num_options).
      normal    : unchanged
      expand    : O text options per image (VCR): image tensors repeated per option, text tensors [B, O, T] -> [B*O, T]
      retrieval : O (caption, image) pairs per example: every tensor [B, O, ...] -> [B*O, ...]
      nlvr      : two images per statement: image tensors [B, 2V, ...] -> [2B, V, ...], each text row duplicated
      dialog    : R rounds x O options (VisDial); as in the reference the mask is expanded from [B, V] straight to four
                  dimensions, which torch rejects - the call fails the same way."""
    batch_size = features.size(0)
    num_options = None
    img, txt = (features, spatials, image_mask), (question, input_mask, segment_ids)
    if process == "dialog":
        nround, num_options = question.size(1), question.size(2)
        target = target.view(-1)
        txt = tuple(_merge_options(_merge_options(t)) for t in txt)
        features, spatials = (_per_option(x, nround * num_options) for x in (features, spatials))
        image_mask = image_mask.unsqueeze(1).expand(batch_size, nround, num_options, image_mask.size(1))   # :172-177
        img = (features, spatials, image_mask.reshape(-1, image_mask.size(-1)))
        batch_size = batch_size * nround
    elif process == "expand":
        num_options = question.size(1)
        img = tuple(_per_option(x, num_options) for x in img)
        txt = tuple(_merge_options(t) for t in txt)
    elif process == "retrieval":
        num_options = question.size(1)
        img = tuple(_merge_options(x) for x in img)
        txt = tuple(_merge_options(t) for t in txt)
    elif process == "nlvr":
        num_options = question.size(1)
        img = tuple(_split_pair(x) for x in img)
        txt = tuple(t.repeat_interleave(2, dim=0) for t in txt)
    return (*img, *txt, target, batch_size, num_options)


def _forward(config, task_cfg, device, task_id, batch, model, criterion, train):
    batch = tuple(t.cuda(device=device, non_blocking=True) for t in batch)
    multi_choice_ids = None
    if task_cfg[task_id]["type"] == "V-logit-mc":
        features, spatials, image_mask, question, target, input_mask, segment_ids, multi_choice_ids, question_id = batch
    else:
        features, spatials, image_mask, question, target, input_mask, segment_ids, question_id = batch
    process = task_cfg[task_id]["process"]
    if process == "dialog" and not train:
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
