"""Stress test of vb_allreduce_f32 (csrc/allreduce.cu) alone, on >= 2 real GPUs under torchrun (pytest wrapper:
tests/test_parallel_gpu.py::test_switch_allreduce_exact).  Integer-valued fp32 data make every sum exact, so the result
must equal the closed form bit for bit - for random bucket offsets / sizes, with and without a multicast address, while a
GEMM-like load runs on the main stream (the situation of the data-parallel backward).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tests/allreduce_check.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def pattern(n, rank, it, device):
    """Integer-valued pseudo-random fp32 data every rank can reproduce for every other rank."""
    idx = torch.arange(n, device=device, dtype=torch.int64)
    return (((idx * 2654435761 + rank * 40503 + it * 9176 + 12345) >> 7) % 2001 - 1000).to(torch.float32)


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    from volta_b200.parallel import SwitchAllReduce
    numel = int(os.environ.get("AR_CHECK_NUMEL", str(48 * 1024 * 1024)))
    iters = int(os.environ.get("AR_CHECK_ITERS", "40"))
    red = SwitchAllReduce(numel, dev)
    modes = ["multicast"] if red.multicast else []
    modes.append("peer")
    mc_ptr = red.multicast
    rnd = random.Random(99)          # the same sequence of buckets on every rank
    load_a = torch.randn(4096, 4096, device=dev, dtype=torch.bfloat16)
    report = {}
    for mode in modes:
        red.multicast = mc_ptr if mode == "multicast" else 0
        worst_bad = 0
        for it in range(iters):
            lo = rnd.randrange(0, numel // 2, 64)
            n = rnd.choice([64, 4096, 1 << 20, 10_000_000 // 64 * 64, rnd.randrange(64, numel // 2, 64)])
            hi = min(numel, lo + n)
            mine = pattern(numel, rank, it, dev)
            red.buf.copy_(mine)
            for _ in range(3):                       # main-stream load next to the collective
                load_a = (load_a @ load_a).clamp_(-1, 1)
            scale = 1.0 / world if it % 2 else 1.0
            red.allreduce(lo, hi, scale)
            for _ in range(3):
                load_a = (load_a @ load_a).clamp_(-1, 1)
            red.wait()
            expect = sum(pattern(numel, r, it, dev) for r in range(world))[lo:hi] * scale
            torch.cuda.synchronize()
            bad = int((red.buf[lo:hi] != expect).sum())
            outside = int((red.buf[:lo] != mine[:lo]).sum()) + int((red.buf[hi:] != mine[hi:]).sum())
            assert bad == 0 and outside == 0, (mode, it, lo, hi, bad, outside)
            worst_bad = max(worst_bad, bad)
        # timing of one 40 MB bucket (1e7 elements), alone
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record(red.stream)
        for _ in range(10):
            red.allreduce(0, 10_000_000 // 64 * 64, 1.0)
        e1.record(red.stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        report[mode] = {"iterations": iters, "mismatches": worst_bad, "ms_per_40MB_bucket": ms,
                        "algbw_GBps": 4e7 / ms / 1e6, "blocks": red.blocks}
    if rank == 0:
        print(json.dumps({"allreduce_check": "ok", "world": world, **report}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
