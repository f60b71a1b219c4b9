#!/bin/bash
# final 8-GPU run: driver-style default bench (all legs), retrieval scoring, multi-rank checks
O=gpurun_out
: > $O/r2_call15.rc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2_b15_8gpu.json 2> $O/r2_b15_8gpu.err; echo "8gpu rc=$?" >> $O/r2_call15.rc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --workload retrieval --steps 10 --warmup 3 > $O/r2_b15_retr8.json 2> $O/r2_b15_retr8.err; echo "retr8 rc=$?" >> $O/r2_call15.rc
DDP_CHECK_ITERS=5 DDP_CHECK_BATCH=8 DDP_CHECK_PRECISION=fp32 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 tests/ddp_check.py > $O/r2_ddp8_fp32.log 2>&1; echo "ddp8 fp32 rc=$?" >> $O/r2_call15.rc
DDP_CHECK_ITERS=6 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29564 tests/ddp_check.py > $O/r2_ddp8_bf16.log 2>&1; echo "ddp8 bf16 rc=$?" >> $O/r2_call15.rc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29565 bench.py --gpus 4 --steps 20 --warmup 5 --no-optimizer-leg --no-full-length-leg > $O/r2_b15_4gpu.json 2> $O/r2_b15_4gpu.err; echo "4gpu rc=$?" >> $O/r2_call15.rc
cat $O/r2_call15.rc; tail -1 $O/r2_ddp8_fp32.log | cut -c1-600; tail -1 $O/r2_ddp8_bf16.log | cut -c1-600
