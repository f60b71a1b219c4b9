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
    rk = RetrievalRanker(10, 2, device="cpu")          # buffers may live anywhere; ranking may not
    with pytest.raises(RuntimeError):
        rk.image_retrieval()
    with pytest.raises(RuntimeError):
        rk.text_retrieval()


class _FakeTasksModel(torch.nn.Module):
    """Stands in for BertForVLTasks in the host-logic test: the 'logit' of a pair is a fixed function of its inputs."""

    def forward(self, question, features, spatials, task, segment_ids, input_mask, image_mask):
        return (features.sum((1, 2)).view(-1, 1) + question[:, :1].float() * 0.001, None, None, None)


def _retrieval_batches(C, I, chunk):
    g = torch.Generator().manual_seed(17)
    feats = torch.randn(I, 3, 4, generator=g)
    out = []
    for c in range(C):
        for part in range(I // chunk):
            sl = slice(part * chunk, (part + 1) * chunk)
            target = (torch.arange(I)[sl] == c % I).long().view(1, chunk)
            out.append((feats[sl].unsqueeze(0), torch.zeros(1, chunk, 3, 5), torch.ones(1, chunk, 3), torch.full((1, 6), c),
                        torch.ones(1, 6), torch.zeros(1, 6), target, torch.tensor(c), torch.tensor(part)))
    return out


def _sharded_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from volta_b200.retrieval import RetrievalRanker, evaluate_retrieval
        C, I, chunk = 9, 8, 4
        batches = _retrieval_batches(C, I, chunk)
        rk = evaluate_retrieval(_FakeTasksModel(), "TASK8", batches, RetrievalRanker(C, I, "cpu", chunk=chunk), "cpu")
        mine = sum(rk.owns(k) for k in range(len(batches)))
        assert mine in (len(batches) // world, len(batches) // world + 1)
        out.put((rank, rk.score_matrix.clone(), rk.target_matrix.clone()))
    except Exception as e:  # noqa
        out.put((rank, repr(e), None))
    finally:
        dist.destroy_process_group()


def test_sharded_evaluation_rebuilds_the_single_process_matrices_gloo():
    """Two gloo ranks score disjoint halves of the (caption x chunk) grid; after merge() both hold exactly the matrices
    one process computes alone."""
    import socket
    import torch.multiprocessing as mp
    from volta_b200.retrieval import RetrievalRanker, evaluate_retrieval
    C, I, chunk = 9, 8, 4
    single = evaluate_retrieval(_FakeTasksModel(), "TASK8", _retrieval_batches(C, I, chunk),
                                RetrievalRanker(C, I, "cpu", chunk=chunk), "cpu")
    assert int(single.target_matrix.sum()) == C
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, scores, targets in results:
        assert targets is not None, scores
        assert torch.equal(scores, single.score_matrix) and torch.equal(targets, single.target_matrix)


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
