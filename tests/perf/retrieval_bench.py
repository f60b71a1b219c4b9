"""Retrieval ranking (vb_retrieval_ranks) on the reference's Flickr30k evaluation layout: 5000 captions x 1000 images.

    python tests/perf/retrieval_bench.py

Lives under tests/ because its CPU-baseline leg times the oracle (test infrastructure).  Prints ONE JSON line: device time of the row pass (+ top-20) and the column pass (CUDA events, 20 launches each;
algorithmic bytes = the 20 MB score matrix + the 5 MB target matrix read once per pass), and the same ranking done the
reference's way - numpy argsort per row / column - by the oracle on one host core (the timed CPU baseline only)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

from volta_b200 import ops
from volta_b200.retrieval import RetrievalRanker
from volta_b200.synthetic import retrieval_case

C, I = 5000, 1000
scores, targets = retrieval_case(C, I, 3)
rk = RetrievalRanker(C, I, "cuda:0")
rk.load(scores, targets)
rank_c = torch.empty(C, dtype=torch.int32, device="cuda:0")
rank_i = torch.empty(I, dtype=torch.int32, device="cuda:0")
top = torch.empty(C, 20, dtype=torch.int32, device="cuda:0")


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


nbytes = C * I * 5
out = {"op": "vb_retrieval_ranks", "workload": "5000 captions x 1000 images (eval_retrieval.py:164-165)", "bytes_per_pass": nbytes}
us = timed(lambda: ops.retrieval_ranks(rk.score_matrix, rk.target_matrix, 0, rank_c, None))
out["rows"] = {"us": us, "gbs": nbytes / us / 1e3}
us = timed(lambda: ops.retrieval_ranks(rk.score_matrix, rk.target_matrix, 0, rank_c, top))
out["rows+top20"] = {"us": us}
us = timed(lambda: ops.retrieval_ranks(rk.score_matrix, rk.target_matrix, 1, rank_i))
out["columns"] = {"us": us, "gbs": nbytes / us / 1e3}

from oracle.retrieval_oracle import image_retrieval_ranks, text_retrieval_ranks  # noqa: E402  (timed CPU baseline only)

s64, t64 = scores.double().numpy(), targets.numpy()
t0 = time.perf_counter()
image_retrieval_ranks(s64, t64)
t1 = time.perf_counter()
text_retrieval_ranks(s64, t64)
t2 = time.perf_counter()
out["cpu_oracle"] = {"rows_ms": (t1 - t0) * 1e3, "columns_ms": (t2 - t1) * 1e3, "cores": 1, "kind": "port"}
print(json.dumps(out))
