#!/bin/bash
# round 2, GPU call 6 (8 GPUs): scaling A/B of the gradient exchange, multi-rank checks, BASELINE configs 3/4/5 at 8 GPUs
O=gpurun_out
run() {  # name, bench args..., env via ENVV
  name=$1; shift
  env $ENVV timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 "$@" > $O/r2_b6_$name.json 2> $O/r2_b6_$name.err; echo "$name rc=$?" >> $O/r2_call6.rc
}
: > $O/r2_call6.rc
ENVV="VOLTA_B200_DDP_ALLREDUCE=switch" run switch64 --steps 20 --warmup 5 --no-cpu-baseline --no-optimizer-leg --no-full-length-leg
ENVV="VOLTA_B200_DDP_ALLREDUCE=switch VOLTA_B200_DDP_BLOCKS=148" run switch148 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg
ENVV="VOLTA_B200_DDP_ALLREDUCE=switch VOLTA_B200_DDP_BLOCKS=32" run switch32 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg
ENVV="VOLTA_B200_DDP_ALLREDUCE=nccl" run nccl --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg
ENVV="VOLTA_B200_DDP_ALLREDUCE=nccl VOLTA_B200_DDP_BF16=1" run nccl_bf16 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg
AR_CHECK_ITERS=12 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tests/allreduce_check.py > $O/r2_ar_check8.log 2>&1; echo "ar_check8 rc=$?" >> $O/r2_call6.rc
DDP_CHECK_ITERS=4 DDP_CHECK_WIRE=fp32 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 tests/ddp_check.py > $O/r2_ddp_switch8.log 2>&1; echo "ddp switch8 rc=$?" >> $O/r2_call6.rc
ENVV="X=1" run lxmert --model ctrl_lxmert --steps 20 --warmup 5 --no-cpu-baseline --no-optimizer-leg --no-full-length-leg
ENVV="X=1" run recipe32 --batch 32 --seq-len 38 --steps 30 --warmup 5 --no-cpu-baseline --no-optimizer-leg --no-full-length-leg
ENVV="X=1" run retrieval --workload retrieval --steps 10 --warmup 3 --no-cpu-baseline
ENVV="X=1" run visualbert --model ctrl_visualbert_base --steps 20 --warmup 5 --no-cpu-baseline --no-optimizer-leg --no-full-length-leg
cat $O/r2_call6.rc; tail -2 $O/r2_ar_check8.log; tail -2 $O/r2_ddp_switch8.log
