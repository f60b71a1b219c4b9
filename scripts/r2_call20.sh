#!/bin/bash
# round 2, call 20 (2 GPUs): multi-rank tests with the last build; the recipe shape (side-stream weight gradients active) under DDP
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -x -q > $O/c20_tests.log 2>&1; echo "tests rc=$?" >> $O/c20_tests.log; tail -3 $O/c20_tests.log
DDP_CHECK_ITERS=5 DDP_CHECK_BATCH=8 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/ddp_check.py > $O/c20_ddp_small.log 2>&1; echo "ddp small rc=$?"; tail -1 $O/c20_ddp_small.log | cut -c1-500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --batch 32 --seq-len 38 --steps 30 --warmup 5 --no-cpu-baseline --no-optimizer-leg --no-full-length-leg > $O/c20_b32_2gpu.json 2> $O/c20_b32_2gpu.err; echo "b32 2gpu rc=$?"
python -c "import json;d=json.load(open('$O/c20_b32_2gpu.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-optimizer-leg --no-full-length-leg > $O/c20_b256_2gpu.json 2> $O/c20_b256_2gpu.err; echo "b256 2gpu rc=$?"
python -c "import json;d=json.load(open('$O/c20_b256_2gpu.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'])"
