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


@pytest.mark.gpu
@pytest.mark.parametrize("wire,precision", [("fp32", "bf16"), ("bf16", "bf16"), ("fp32", "fp32")])
def test_ddp_gradients_equal_mean_of_single_rank_gradients(wire, precision):
    """>= 2 real ranks over NCCL (self-skips on a one-GPU box): tests/ddp_check.py under torchrun - 10 iterations of the
    graph-replayed, bucket-overlapped backward; every parameter gradient equals the mean of the single-rank gradients
    (apex/tests/distributed/DDP/ddp_race_condition_test.py:29-69 is the pattern)."""
    import json
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DDP_CHECK_WIRE=wire, DDP_CHECK_PRECISION=precision)
    if precision == "fp32":    # SIMT validation kernels: a smaller batch and fewer iterations keep the run short
        env.update(DDP_CHECK_BATCH="8", DDP_CHECK_ITERS="5")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "ddp_check.py")],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    rep = json.loads(line)
    assert rep["ddp_check"] == "ok" and rep["buckets_per_step"] >= 2 and rep["backward_segments"] >= 2, rep
    print(line)


@pytest.mark.gpu
def test_switch_allreduce_exact():
    """vb_allreduce_f32 alone on >= 2 real GPUs: tests/allreduce_check.py under torchrun - integer-valued data, random
    bucket offsets / sizes, multicast and peer-pointer modes, a matmul load on the main stream; every result exact."""
    import json
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "allreduce_check.py")],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["allreduce_check"] == "ok"
    print(line)
