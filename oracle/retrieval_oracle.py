"""CPU oracle for the retrieval ranking of eval_retrieval.py.  TEST INFRASTRUCTURE ONLY.

numpy restatement of the ranking the reference computes on the host from its score / target matrices:
  * image retrieval   eval_retrieval.py:201-208 (position of the first matching image in argsort(-row)), top-20 :222
  * text retrieval    eval_retrieval.py:248-256 (best position of a matching caption in argsort(-column))
  * recall metrics    eval_retrieval.py:225-230 / :258-263 (r1, r5, r10, median rank, mean rank)
Restated with a STABLE argsort: the reference's `np.argsort` (quicksort) leaves the order of equal scores unspecified.

Parity status: PINNED for distinct scores - `tests/golden/make_golden_retrieval.py` executes the reference's own
statements (read from /root/reference/eval_retrieval.py, lines cited above) on seeded matrices and commits the ranks
under `tests/golden/retrieval_ranks.pt`; `tests/test_retrieval.py` checks this file against them exactly.

Nothing in `volta_b200/` imports this module; only `tests/` (incl. the timed CPU-baseline leg of `tests/perf/`) and
`__graft_entry__.smoke()` may.
"""
import numpy as np


def image_retrieval_ranks(score_matrix, target_matrix, topk=20):
    C, I = score_matrix.shape
    ranks = np.full((C,), I, np.int64)
    top = np.full((C, topk), -1, np.int64)
    for c in range(C):
        order = np.argsort(-score_matrix[c], kind="stable")
        top[c, :min(topk, I)] = order[:topk]
        tgt = np.where(target_matrix[c] == 1)[0]
        if len(tgt):
            ranks[c] = np.where(order == tgt[0])[0][0]
    return ranks, top


def text_retrieval_ranks(score_matrix, target_matrix):
    C, I = score_matrix.shape
    ranks = np.full((I,), C, np.int64)
    for i in range(I):
        tgt = np.where(target_matrix[:, i] == 1)[0]
        order = np.argsort(-score_matrix[:, i], kind="stable")
        if len(tgt):
            ranks[i] = min(np.where(order == t)[0][0] for t in tgt)
    return ranks


def recall_metrics(rank_matrix):
    rank_matrix = np.asarray(rank_matrix, dtype=np.float64)
    r1 = 100.0 * np.sum(rank_matrix < 1) / len(rank_matrix)
    r5 = 100.0 * np.sum(rank_matrix < 5) / len(rank_matrix)
    r10 = 100.0 * np.sum(rank_matrix < 10) / len(rank_matrix)
    medr = np.floor(np.median(rank_matrix) + 1)
    meanr = np.mean(rank_matrix) + 1
    return dict(r1=float(r1), r5=float(r5), r10=float(r10), medr=float(medr), meanr=float(meanr))
