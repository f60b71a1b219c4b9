#!/bin/bash
# round 2, GPU call 5 (2 GPUs): all-reduce kernel stress test, DDP check (switch + nccl), 1-GPU grouped-wgrad A/B, 2-GPU A/B, GEMM tile sweep
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tests/allreduce_check.py > $O/r2_ar_check.log 2>&1; echo "ar_check rc=$?" > $O/r2_call5.rc
DDP_CHECK_WIRE=fp32 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 tests/ddp_check.py > $O/r2_ddp_switch.log 2>&1; echo "ddp switch rc=$?" >> $O/r2_call5.rc
VOLTA_B200_DDP_ALLREDUCE=nccl DDP_CHECK_WIRE=fp32 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546 tests/ddp_check.py > $O/r2_ddp_nccl.log 2>&1; echo "ddp nccl rc=$?" >> $O/r2_call5.rc
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg > $O/r2_b5_1gpu.json 2> $O/r2_b5_1gpu.err; echo "1gpu rc=$?" >> $O/r2_call5.rc
VOLTA_B200_WGRAD_GROUP=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg > $O/r2_b5_1gpu_nogroup.json 2> $O/r2_b5_1gpu_nogroup.err; echo "1gpu nogroup rc=$?" >> $O/r2_call5.rc
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg > $O/r2_b5_$name.json 2> $O/r2_b5_$name.err; echo "$name rc=$?" >> $O/r2_call5.rc
}
run switch148 VOLTA_B200_DDP_ALLREDUCE=switch
run switch296 VOLTA_B200_DDP_ALLREDUCE=switch VOLTA_B200_DDP_BLOCKS=296
run switch64 VOLTA_B200_DDP_ALLREDUCE=switch VOLTA_B200_DDP_BLOCKS=64
run nccl VOLTA_B200_DDP_ALLREDUCE=nccl
timeout 300 python scripts/gemm_sweep.py > $O/r2_gemm_sweep.log 2>&1; echo "sweep rc=$?" >> $O/r2_call5.rc
cat $O/r2_call5.rc; tail -3 $O/r2_ar_check.log; tail -3 $O/r2_ddp_switch.log; tail -3 $O/r2_ddp_nccl.log
