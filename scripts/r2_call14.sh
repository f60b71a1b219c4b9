#!/bin/bash
# driver-style commands at 2 GPUs (all legs), the multi-rank tests, and the reference arm under torchrun
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2_b14_2gpu.json 2> $O/r2_b14_2gpu.err; echo "2gpu full rc=$?" > $O/r2_call14.rc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/r2_b14_ref2.json 2> $O/r2_b14_ref2.err; echo "ref2 rc=$?" >> $O/r2_call14.rc
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -q -s > $O/r2_tests14.log 2>&1; echo "ddp tests rc=$?" >> $O/r2_call14.rc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --workload retrieval --steps 10 --warmup 3 > $O/r2_b14_retr2.json 2> $O/r2_b14_retr2.err; echo "retr2 rc=$?" >> $O/r2_call14.rc
cat $O/r2_call14.rc; tail -3 $O/r2_tests14.log; cat $O/r2_b14_2gpu.json | head -c 600
