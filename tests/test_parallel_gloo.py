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


# ---------------------------------------------------------------------------------------------------------------------
# The wrapper itself: volta_b200.parallel.DistributedDataParallel around a model whose C ABI is emulated in PyTorch
# (tests/emu_ops.py) - parameter broadcast, progress hooks of the hand-written backward, bucket cuts, final flush.
def _ddp_build():
    import volta_b200
    from golden_util import load_case
    from volta_b200.synthetic import seeded_state_dict
    g, cfg, _ = load_case("mini_ctrl_vilbert")
    model = volta_b200.BertForVLPreTraining(cfg)
    model.load_state_dict(seeded_state_dict(model.state_dict(), seed=0))
    model.set_precision("fp32").eval()
    return cfg, model


def _ddp_step(model, cfg, seed):
    from volta_b200.synthetic import pretraining_batch
    b = pretraining_batch(4, 8, 9, vocab_size=cfg.vocab_size, feat_dim=cfg.v_feature_size, seed=seed)
    model.zero_grad(set_to_none=True)
    l = model(b["input_ids"], b["image_feat"], b["image_loc"], b["segment_ids"], b["input_mask"], b["image_mask"],
              b["lm_label_ids"], b["image_label"], b["image_cls"], None, None, None, None, None, b["is_match"])
    (l[0] + l[1] + l[2]).sum().backward()
    return {k: p.grad.clone() for k, p in model.named_parameters()}


def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import emu_ops
        emu_ops.install(pytest.MonkeyPatch())
        from volta_b200.parallel import DistributedDataParallel
        cfg, ref_model = _ddp_build()
        singles = [_ddp_step(ref_model, cfg, 100 + r) for r in range(world)]      # what each rank computes alone
        report = {}
        for wire in (None, torch.bfloat16):
            cfg, model = _ddp_build()
            if rank == 1:
                with torch.no_grad():
                    for p in model.parameters():
                        p.add_(1.0)                     # the broadcast from rank 0 must overwrite this
            ddp = DistributedDataParallel(model, message_size=200_000, allreduce_dtype=wire)
            grads = _ddp_step(ddp, cfg, 100 + rank)
            worst = 0.0
            for k, got in grads.items():
                expect = sum(s[k[len("module."):]] for s in singles) / world
                worst = max(worst, float((got - expect).norm()) / max(float(expect.norm()), 1e-9))
            report[str(wire)] = (worst, ddp.last_num_buckets)
        out.put((rank, report))
    except Exception:  # noqa
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_ddp_wrapper_over_the_emulated_engine_two_ranks_gloo():
    """Every parameter gradient after backward() through the wrapper equals the mean of the two ranks' own gradients
    (exactly with the fp32 wire, to bf16 rounding with allreduce_dtype=torch.bfloat16); the backward closed several
    buckets on the way (apex: delay_allreduce=False, message_size)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, report in results:
        assert isinstance(report, dict), report
        worst, buckets = report["None"]
        assert worst < 1e-6 and buckets > 2, (rank, report)
        worst16, buckets16 = report["torch.bfloat16"]
        assert worst16 < 1e-2 and buckets16 == buckets, (rank, report)


def test_bucketer_tiles_the_buffer_for_any_frontier_sequence():
    """Property (seeded random cases, one-rank gloo group): whatever the boundaries, message size and order of frontier
    notifications, the issued ranges tile [0, total) exactly once from the end of the buffer towards offset 0, every
    range but the flush reaches the message size, and simulate() - what the CUDA-graph executor cuts the backward by -
    predicts exactly the notifications that issue a collective."""
    import random
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), RANK="0", WORLD_SIZE="1")
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        from volta_b200.parallel import GradBucketer
        rnd = random.Random(1234)
        for case in range(60):
            total = rnd.randint(50, 5000)
            bounds = sorted(set(rnd.randint(0, total) for _ in range(rnd.randint(0, 25))))
            msg = rnd.choice([1, 10, 100, 1000, 10 ** 7])
            flat = torch.arange(total, dtype=torch.float32)
            b = GradBucketer(flat, bounds, message_size=msg, wire_dtype=rnd.choice([None, torch.bfloat16]))
            cut = b.simulate()
            frontiers = sorted((rnd.randint(0, total) for _ in range(rnd.randint(1, 30))), reverse=True)
            if rnd.random() < 0.5:
                frontiers.append(0)
            for f in frontiers:
                before = len(b.ranges)
                predicted = cut(f)
                b.advance(f)
                assert predicted == (len(b.ranges) > before), (case, f)
                assert all(lo >= f for lo, hi in b.ranges), (case, f, b.ranges)   # nothing below the frontier is sent
            n = b.finish()
            assert n == len(b.ranges)
            hi = total
            for lo, h in b.ranges:
                assert h == hi and lo < h, (case, b.ranges)
                hi = lo
            assert hi == 0, (case, b.ranges)
            for lo, h in b.ranges[:-1]:
                assert h - lo >= msg or lo == 0, (case, msg, b.ranges)
            assert torch.equal(flat, torch.arange(total, dtype=torch.float32).bfloat16().float() if b.wire is not None
                               else torch.arange(total, dtype=torch.float32))
    finally:
        dist.destroy_process_group()
