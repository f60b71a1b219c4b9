"""Bucketed gradient all-reduce on the device with a one-rank NCCL group: the fp32 wire leaves the buffer untouched, the
bf16 wire (allreduce_dtype=torch.bfloat16) rounds every element once through vb_eltwise and writes fp32 back."""
import os
import socket

import pytest
import torch
import torch.distributed as dist


@pytest.mark.gpu
def test_bucketer_wire_formats_on_one_rank_nccl():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        from volta_b200.parallel import GradBucketer
        total = 100_003                                  # odd length: the last bucket takes the scalar kernel
        bounds = [0, 1000, 25_000, 40_008, 64_000, 90_001]
        g = torch.Generator(device="cuda").manual_seed(3)
        base = torch.randn(total, device="cuda", generator=g)
        for wire in (None, torch.bfloat16):
            flat = base.clone()
            b = GradBucketer(flat, bounds, message_size=20_000, wire_dtype=wire)
            for frontier in (90_001, 64_000, 40_008, 25_000, 1000):
                b.advance(frontier)
            n = b.finish()
            torch.cuda.synchronize()
            assert n >= 3 and sorted(b.ranges)[0][0] == 0 and sorted(b.ranges)[-1][1] == total
            expect = base if wire is None else base.bfloat16().float()
            assert torch.equal(flat, expect)
    finally:
        dist.destroy_process_group()
