"""Retrieval evaluation with the score matrix kept and ranked on the device.

Mirrors the bookkeeping of the reference's `eval_retrieval.py` main loop: a `[num_captions, num_images]` score matrix
and target matrix filled chunk by chunk (`:164-165, 183-198` - 500 candidate images per model call), the rank of the
matching image per caption (`:201-208`), the top-20 list per caption (`:222`), recall@1/5/10 + median / mean rank
(`:225-230`), and the text-retrieval pass over the columns (`:248-263`).  The reference copies every chunk of logits to
the host and argsorts rows / columns in numpy; here the logits stay in HBM and `vb_retrieval_ranks` turns the
matrices into ranks (a rank is a count, no sort).  Only the ranks (4 B per caption / image) cross to the host, where
the five summary numbers are formed with the reference's own formulas.

Ties between scores keep index order (the reference's unstable argsort leaves them unspecified).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import ops


def recall_metrics(ranks):
    """r1, r5, r10, medr, meanr from 0-based ranks (eval_retrieval.py:225-230 / :258-263)."""
    r = np.asarray(ranks, dtype=np.float64)
    n = len(r)
    return dict(r1=100.0 * float(np.sum(r < 1)) / n, r5=100.0 * float(np.sum(r < 5)) / n,
                r10=100.0 * float(np.sum(r < 10)) / n, medr=float(np.floor(np.median(r) + 1)), meanr=float(np.mean(r) + 1))


class RetrievalRanker:
    """ranker = RetrievalRanker(5000, 1000, device)
    for each batch:  ranker.update(caption_idx, image_idx, vil_logit, target)     # 500 candidates per call
    ranks, top20, metrics = ranker.image_retrieval();  ranks_t, metrics_t = ranker.text_retrieval()"""

    def __init__(self, num_captions=5000, num_images=1000, device="cuda", chunk=500, topk=20):
        # the matrices are buffers (any device: the bookkeeping and the cross-rank merge are plumbing); RANKING has no
        # CPU path - image_retrieval() / text_retrieval() raise unless the buffers live on a CUDA device
        device = torch.device(device)
        self.C, self.I, self.chunk, self.topk = int(num_captions), int(num_images), int(chunk), int(topk)
        self.score_matrix = torch.zeros(self.C, self.I, dtype=torch.float32, device=device)
        self.target_matrix = torch.zeros(self.C, self.I, dtype=torch.uint8, device=device)

    def update(self, caption_idx, image_idx, vil_logit, target):
        """score_matrix[caption_idx, image_idx*chunk : (image_idx+1)*chunk] = vil_logit.view(-1), same for the targets
        (eval_retrieval.py:191-198); device-to-device, no host copy."""
        c, lo = int(caption_idx), int(image_idx) * self.chunk
        n = vil_logit.numel()
        self.score_matrix[c, lo:lo + n].copy_(vil_logit.reshape(-1))
        self.target_matrix[c, lo:lo + n].copy_(target.reshape(-1) == 1)

    def load(self, scores, targets):
        """Whole matrices at once (tests, or scores gathered from the ranks of a sharded evaluation)."""
        self.score_matrix.copy_(scores.to(self.score_matrix.device, non_blocking=True))
        self.target_matrix.copy_((targets == 1).to(self.target_matrix.device, non_blocking=True))

    # ---- data parallel evaluation (retrieval row) -------------------------------------------------
    def owns(self, step, rank=None, world=None):
        """Work item `step` (the i-th batch of the evaluation loop, eval_retrieval.py:168: one caption against one chunk
        of 500 images) belongs to rank step % world.  The reference runs the loop on ONE GPU (:150-152 rejects more);
        here every rank scores its share of the (caption x chunk) grid with its own model replica - no collective on the
        data path."""
        if world is None:
            world = dist.get_world_size() if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        return int(step) % world == rank

    def merge(self, group=None):
        """One exchange at the end: every pair was scored by exactly one rank and is zero elsewhere, so a SUM of the
        score matrices (x + 0 is exact) and a MAX of the target matrices reproduce the single-GPU matrices on every
        rank (20 MB + 5 MB for Flickr30k)."""
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.score_matrix, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.target_matrix, op=dist.ReduceOp.MAX, group=group)

    def _require_device(self):
        if not self.score_matrix.is_cuda:
            raise RuntimeError("volta_b200.retrieval has no CPU path: ranking needs the matrices on a CUDA (sm_100a) device")

    def image_retrieval(self, num_captions=None):
        """Rank of the first matching image for captions [0, num_captions) (the running print of :210-220 uses a prefix)."""
        self._require_device()
        C = self.C if num_captions is None else int(num_captions)
        rank = torch.empty(C, dtype=torch.int32, device=self.score_matrix.device)
        top = torch.empty(C, self.topk, dtype=torch.int32, device=self.score_matrix.device)
        ops.retrieval_ranks(self.score_matrix[:C], self.target_matrix[:C], 0, rank, top)
        ranks = rank.cpu().numpy()
        return ranks, top.cpu().numpy(), recall_metrics(ranks)

    def text_retrieval(self):
        self._require_device()
        rank = torch.empty(self.I, dtype=torch.int32, device=self.score_matrix.device)
        ops.retrieval_ranks(self.score_matrix, self.target_matrix, 1, rank)
        ranks = rank.cpu().numpy()
        return ranks, recall_metrics(ranks)


class GraphedScorer:
    """CUDA-graph replay of the forward-only scoring call `model(question, features, spatials, task, segment_ids,
    input_mask, image_mask)[0]` for a fixed input signature: the ~280 kernel launches of one call cost the host ~8 ms
    when issued one by one (more with eight ranks sharing the host's cores), a replay costs microseconds.  First call with
    a signature: eager; second: capture; then replays.  The returned logits live in a static buffer that the next replay
    overwrites - the caller consumes them (RetrievalRanker.update copies them out) before scoring the next item."""

    def __init__(self, model, task):
        self.model, self.task, self.cache = model, task, {}

    def _call(self, question, features, spatials, segment_ids, input_mask, image_mask):
        return self.model(question, features, spatials, self.task, segment_ids, input_mask, image_mask)[0]

    def __call__(self, *tensors):
        import gc
        import os
        from . import engine
        if os.environ.get("VOLTA_B200_GRAPHS", "1") == "0" or self.model.training or not tensors[0].is_cuda:
            return self._call(*tensors)
        key = tuple((tuple(t.shape), t.dtype) for t in tensors)
        ent = self.cache.get(key)
        if ent is None:
            self.cache[key] = "seen"
            return self._call(*tensors)
        store = engine.model_store(self.model)
        store.ensure(tensors[0].device, self.model.precision)     # the bf16 operand copy is fresh before capture / replay
        ptrs = (store.flat32.data_ptr(), None if store.flat16 is None else store.flat16.data_ptr())
        if ent == "seen" or ent[3] != ptrs:
            static_in = [t.clone() for t in tensors]
            gc.collect()
            was = gc.isenabled()
            gc.disable()
            n0 = engine.launch_count
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    out = self._call(*static_in)
            finally:
                if was:
                    gc.enable()
            n = engine.launch_count - n0
            engine.launch_count -= n        # a capture launches nothing; replays are counted
            ent = self.cache[key] = (g, static_in, out, ptrs, n)
        g, static_in, out, _, n = ent
        for dst, src in zip(static_in, tensors):
            dst.copy_(src, non_blocking=True)
        g.replay()
        engine.launch_count += n
        return out


def _prefetched(items, device):
    """(caption_idx, image_idx, device tensors) of every work item, with the host->device copy of item i+1 issued on a
    copy stream while item i is being scored (pinned host tensors travel asynchronously; the reference's loop calls
    `.cuda(non_blocking=True)` on the compute stream, eval_retrieval.py:170)."""
    device = torch.device(device)
    if device.type != "cuda":
        for b in items:
            yield int(b[7]), int(b[8]), tuple(t.to(device) for t in b[:7])
        return
    copy = torch.cuda.Stream(device=device)

    def issue(b):
        # the two indices are host integers in the reference's loader: read them before anything is queued, so that
        # the loop never waits for the device
        ci, ii = int(b[7]), int(b[8])
        with torch.cuda.stream(copy):
            dev = tuple(t.to(device, non_blocking=True) for t in b[:7])
            ev = torch.cuda.Event()
            ev.record(copy)
        return ci, ii, dev, ev

    it = iter(items)
    first = next(it, None)
    cur = issue(first) if first is not None else None
    while cur is not None:
        nxt_b = next(it, None)
        nxt = issue(nxt_b) if nxt_b is not None else None
        ci, ii, dev, ev = cur
        main = torch.cuda.current_stream(device)
        main.wait_event(ev)
        for t in dev:
            t.record_stream(main)
        yield ci, ii, dev
        cur = nxt


def evaluate_retrieval(model, task, batches, ranker, device, group=None, start=0, merge=True):
    """The evaluation loop of eval_retrieval.py:168-198 over `batches` (tuples as RetrievalDatasetVal yields them:
    features, spatials, image_mask, question, input_mask, segment_ids, target, caption_idx, image_idx), sharded over the
    ranks of `group` by batch index, merged once at the end.  `model` is a BertForVLTasks replica in eval mode; returns
    the ranker (matrices complete on every rank).  `start`: index of the first batch in the whole evaluation (a resumed or
    chunked loop); `merge=False` leaves the cross-rank merge to a later `ranker.merge()`."""
    scorer = model.__dict__.setdefault("_graphed_scorers", {}).setdefault(task, None) or GraphedScorer(model, task)
    model.__dict__["_graphed_scorers"][task] = scorer
    rank = None if group is None else dist.get_rank(group)
    world = None if group is None else dist.get_world_size(group)
    owned = (batch for step, batch in enumerate(batches, start) if ranker.owns(step, rank, world))
    with torch.no_grad():
        for caption_idx, image_idx, dev in _prefetched(owned, device):
            features, spatials, image_mask, question, input_mask, segment_ids, target = dev
            features, spatials, image_mask = features.squeeze(0), spatials.squeeze(0), image_mask.squeeze(0)
            n = features.size(0)
            question, segment_ids, input_mask = question.repeat(n, 1), segment_ids.repeat(n, 1), input_mask.repeat(n, 1)
            vil_logit = scorer(question, features, spatials, segment_ids, input_mask, image_mask)
            ranker.update(caption_idx, image_idx, vil_logit, target)
    if merge:
        ranker.merge(group)
    return ranker
