"""Golden ranks for the retrieval ranking, produced by the reference's OWN statements (build container only).

    python tests/golden/make_golden_retrieval.py      # writes tests/golden/retrieval_ranks.pt

`eval_retrieval.py` is a script (the ranking lives inside `main()` next to argument parsing, data loading and the
model), so it cannot be imported; instead the statements that compute the ranks are read from the file by line number
- image retrieval :201-208, top-20 list :222, text retrieval :249-256, metrics :225-230 and :258-263 - dedented, and executed
unchanged on seeded score / target matrices of the reference's Flickr30k layout (5 captions per image, every caption
scored against every image).  Scores are distinct (ties are unspecified under the reference's unstable argsort).
Inputs are regenerated from the seed by `retrieval_case`; stored are the reference's ranks, top-20 lists and metrics.
"""
import os
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from volta_b200.synthetic import retrieval_case  # noqa: E402

CASES = {"small": (60, 12, 1), "wide": (35, 7, 2), "flickr": (5000, 1000, 3)}


def ref_lines(lo, hi):
    src = open("/root/reference/eval_retrieval.py").read().split("\n")
    return textwrap.dedent("\n".join(src[lo - 1:hi]))


def main():
    rank_stmt = ref_lines(201, 208)
    top_stmt = ref_lines(222, 222)
    text_block = ref_lines(249, 263).replace("range(1000)", "range(score_matrix.shape[1])").replace(
        "np.zeros(1000)", "np.zeros(score_matrix.shape[1])")
    metric_block = ref_lines(225, 230)
    assert rank_stmt.startswith("rank = np.where(") and "results.append" in top_stmt and "tgt_captions" in text_block
    out = {}
    for name, (C, I, seed) in CASES.items():
        scores, targets = retrieval_case(C, I, seed)
        score_matrix = scores.double().numpy()      # the reference stores fp32 logits into float64 matrices (:164-165)
        target_matrix = targets.double().numpy()
        ns = dict(np=np, score_matrix=score_matrix, target_matrix=target_matrix, rank_matrix=np.ones(C) * I, results=[])
        for c in range(C):
            ns["caption_idx"] = c
            exec(rank_stmt, ns)
            exec(top_stmt, ns)
        image_ranks = ns["rank_matrix"].copy()
        exec(metric_block, ns)
        image_metrics = {k: float(ns[k]) for k in ("r1", "r5", "r10", "medr", "meanr")}
        exec(text_block, ns)
        text_ranks = ns["rank_matrix"].copy()
        text_metrics = {k: float(ns[k]) for k in ("r1", "r5", "r10", "medr", "meanr")}
        out[name] = dict(case=(C, I, seed), image_ranks=torch.from_numpy(image_ranks).long(),
                         top20=torch.tensor([r + [-1] * (20 - len(r)) for r in ns["results"]]).to(torch.int16),
                         text_ranks=torch.from_numpy(text_ranks).long(), image_metrics=image_metrics, text_metrics=text_metrics)
        print(name, image_metrics, text_metrics)
    torch.save(out, os.path.join(HERE, "retrieval_ranks.pt"))


if __name__ == "__main__":
    main()
