"""Data-parallel host logic on CPU: two gloo ranks run the bucketed gradient all-reduce
(volta_b200.parallel.GradBucketer) over a flat buffer, following the pattern of the reference's DDP race test
(apex/tests/distributed/DDP/ddp_race_condition_test.py:29-69: closed-form expected sums)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from volta_b200.parallel import GradBucketer
        total = 1000
        bounds = [0, 100, 250, 400, 640, 900]
        for it in range(3):
            flat = torch.full((total,), float(rank + 1 + it))
            b = GradBucketer(flat, bounds, message_size=200)
            # the backward retires parameters from the end of the buffer towards offset 0
            for frontier in (900, 640, 400, 250, 100):
                b.advance(frontier)
            n = b.finish()
            expect = sum(r + 1 + it for r in range(world)) / world
            assert torch.allclose(flat, torch.full((total,), expect)), (rank, it)
            covered = sorted(b.ranges)
            assert covered[0][0] == 0 and covered[-1][1] == total
            for (lo0, hi0), (lo1, hi1) in zip(covered, covered[1:]):
                assert hi0 == lo1
            # 640..1000 (360 >= 200) closes first, then 400..640 (240), then 100..400 (300); finish() flushes 0..100
            assert b.ranges == [(640, 1000), (400, 640), (100, 400), (0, 100)], b.ranges
            assert n == 4
        # bf16 wire (allreduce_dtype=torch.bfloat16 / VOLTA_B200_DDP_BF16=1): same buckets, values rounded to bf16 once
        # before the sum; the fp32 buffer receives the averaged result
        gen = torch.Generator().manual_seed(7)
        base = torch.randn(total, generator=gen)
        flat = base * (rank + 1)
        b = GradBucketer(flat, bounds, message_size=200, wire_dtype=torch.bfloat16)
        for frontier in (900, 640, 400, 250, 100):
            b.advance(frontier)
        assert b.finish() == 4 and b.ranges == [(640, 1000), (400, 640), (100, 400), (0, 100)]
        expect = sum((base * (r + 1)).bfloat16().float() for r in range(world)) / world
        assert flat.dtype == torch.float32
        assert torch.allclose(flat, expect, rtol=2 ** -7, atol=1e-6), (flat - expect).abs().max()
        exact = base * sum(r + 1 for r in range(world)) / world
        assert torch.allclose(flat, exact, rtol=2 ** -6, atol=1e-6)
        out.put((rank, "ok"))
    except Exception as e:  # noqa
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
