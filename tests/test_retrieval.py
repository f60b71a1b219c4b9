"""Retrieval ranking: the numpy oracle against the ranks the reference's own statements produce
(tests/golden/retrieval_ranks.pt, made by tests/golden/make_golden_retrieval.py), exactly; the CUDA kernel against the
golden ranks and the oracle (`-m gpu`), including ties, captions without a match and the 5000 x 1000 Flickr30k layout."""
import os

import numpy as np
import pytest
import torch

from oracle.retrieval_oracle import image_retrieval_ranks, recall_metrics, text_retrieval_ranks
from volta_b200.synthetic import retrieval_case

CASES = ["small", "wide", "flickr"]


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "retrieval_ranks.pt"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_equals_the_reference_statements(name, golden):
    g = golden[name]
    C, I, seed = g["case"]
    scores, targets = retrieval_case(C, I, seed)
    ranks, top = image_retrieval_ranks(scores.double().numpy(), targets.numpy(), 20)
    assert np.array_equal(ranks, g["image_ranks"].numpy())
    assert np.array_equal(top, g["top20"].long().numpy())
    assert np.array_equal(text_retrieval_ranks(scores.double().numpy(), targets.numpy()), g["text_ranks"].numpy())
    assert recall_metrics(ranks) == g["image_metrics"]
    assert recall_metrics(g["text_ranks"].numpy()) == g["text_metrics"]


def test_host_metrics_equal_the_reference_formulas(golden):
    from volta_b200.retrieval import recall_metrics as mine
    for name in CASES:
        assert mine(golden[name]["image_ranks"].numpy()) == golden[name]["image_metrics"]
        assert mine(golden[name]["text_ranks"].numpy()) == golden[name]["text_metrics"]


def test_ranker_refuses_to_run_without_a_device():
    from volta_b200.retrieval import RetrievalRanker
    with pytest.raises(RuntimeError):
        RetrievalRanker(10, 2, device="cpu")


# ----------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernel_equals_the_reference_ranks(name, golden):
    from volta_b200.retrieval import RetrievalRanker
    g = golden[name]
    C, I, seed = g["case"]
    scores, targets = retrieval_case(C, I, seed)
    rk = RetrievalRanker(C, I, "cuda:0")
    rk.load(scores, targets)
    ranks, top, metrics = rk.image_retrieval()
    assert np.array_equal(ranks, g["image_ranks"].numpy())
    assert np.array_equal(top, g["top20"].long().numpy())
    assert metrics == g["image_metrics"]
    ranks_t, metrics_t = rk.text_retrieval()
    assert np.array_equal(ranks_t, g["text_ranks"].numpy())
    assert metrics_t == g["text_metrics"]


@pytest.mark.gpu
def test_chunked_updates_ties_and_missing_targets():
    from volta_b200.retrieval import RetrievalRanker
    C, I = 37, 24
    g = torch.Generator().manual_seed(5)
    scores = torch.randint(0, 6, (C, I), generator=g).float()         # many ties
    targets = (torch.rand(C, I, generator=g) < 0.1).to(torch.uint8)     # several matches per row, some rows / columns none
    targets[3] = 0
    targets[:, 7] = 0
    rk = RetrievalRanker(C, I, "cuda:0", chunk=12, topk=30)            # k > I: padded with -1
    for c in range(C):
        for part in range(2):
            rk.update(torch.tensor(c), torch.tensor(part), scores[c, part * 12:(part + 1) * 12].cuda().view(12, 1),
                      targets[c, part * 12:(part + 1) * 12].cuda().view(12, 1))
    ranks, top, _ = rk.image_retrieval()
    o_ranks, o_top = image_retrieval_ranks(scores.double().numpy(), targets.numpy(), 30)
    assert np.array_equal(ranks, o_ranks) and ranks[3] == I
    assert np.array_equal(top, o_top)
    ranks_t, _ = rk.text_retrieval()
    o_t = text_retrieval_ranks(scores.double().numpy(), targets.numpy())
    assert np.array_equal(ranks_t, o_t) and ranks_t[7] == C
    prefix, _, _ = rk.image_retrieval(num_captions=10)
    assert np.array_equal(prefix, o_ranks[:10])
