#!/bin/bash
# round 2, call 26 (2 GPUs): last build under the data-parallel wrapper (bench + gradient check)
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-optimizer-leg --no-full-length-leg > $O/c26_2gpu.json 2> $O/c26_2gpu.err; echo "2gpu rc=$? lines=$(wc -l < $O/c26_2gpu.json)"
python -c "import json;d=json.loads(open('$O/c26_2gpu.json').read());print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
DDP_CHECK_ITERS=4 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29592 tests/ddp_check.py > $O/c26_ddp.log 2>&1; echo "ddp rc=$?"; tail -1 $O/c26_ddp.log | cut -c1-400
