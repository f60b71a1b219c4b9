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
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("volta_b200.retrieval has no CPU path: a CUDA (sm_100a) device is required")
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

    def image_retrieval(self, num_captions=None):
        """Rank of the first matching image for captions [0, num_captions) (the running print of :210-220 uses a prefix)."""
        C = self.C if num_captions is None else int(num_captions)
        rank = torch.empty(C, dtype=torch.int32, device=self.score_matrix.device)
        top = torch.empty(C, self.topk, dtype=torch.int32, device=self.score_matrix.device)
        ops.retrieval_ranks(self.score_matrix[:C], self.target_matrix[:C], 0, rank, top)
        ranks = rank.cpu().numpy()
        return ranks, top.cpu().numpy(), recall_metrics(ranks)

    def text_retrieval(self):
        rank = torch.empty(self.I, dtype=torch.int32, device=self.score_matrix.device)
        ops.retrieval_ranks(self.score_matrix, self.target_matrix, 1, rank)
        ranks = rank.cpu().numpy()
        return ranks, recall_metrics(ranks)
