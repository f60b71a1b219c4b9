#!/bin/bash
# round 2, call 22 (2 GPUs): driver-style torchrun launches after the stdout change (own arm and reference arm)
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 5 > $O/c22_2gpu.json 2> $O/c22_2gpu.err; echo "2gpu rc=$? lines=$(wc -l < $O/c22_2gpu.json)"
python -c "import json;d=json.loads(open('$O/c22_2gpu.json').read());print(d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/c22_ref2.json 2> $O/c22_ref2.err; echo "ref2 rc=$? lines=$(wc -l < $O/c22_ref2.json)"
python -c "import json;d=json.loads(open('$O/c22_ref2.json').read());print(d['impl'], d['value'], d['n_gpus'])"
