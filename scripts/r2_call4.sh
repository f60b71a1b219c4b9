#!/bin/bash
# round 2, GPU call 4 (2 GPUs): whole GPU suite (multi-rank DDP check included), 1-GPU bench (grouped wgrad, pre-zeroed
# gradients), 2-GPU bench: own NVSwitch all-reduce vs NCCL
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > $O/r2_tests4.log 2>&1; echo "tests rc=$?" > $O/r2_call4.rc
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_b4_1gpu.json 2> $O/r2_b4_1gpu.err; echo "1gpu rc=$?" >> $O/r2_call4.rc
VOLTA_B200_WGRAD_GROUP=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg > $O/r2_b4_1gpu_nogroup.json 2> $O/r2_b4_1gpu_nogroup.err; echo "1gpu nogroup rc=$?" >> $O/r2_call4.rc
VOLTA_B200_PREZERO_GRADS=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg > $O/r2_b4_1gpu_noprezero.json 2> $O/r2_b4_1gpu_noprezero.err; echo "1gpu noprezero rc=$?" >> $O/r2_call4.rc
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-optimizer-leg --no-full-length-leg > $O/r2_b4_$name.json 2> $O/r2_b4_$name.err; echo "$name rc=$?" >> $O/r2_call4.rc
}
run switch VOLTA_B200_DDP_ALLREDUCE=switch
run switch_p2p VOLTA_B200_DDP_ALLREDUCE=switch VOLTA_B200_DDP_MULTICAST=0
run switch_b32 VOLTA_B200_DDP_ALLREDUCE=switch VOLTA_B200_DDP_BLOCKS=32
run nccl VOLTA_B200_DDP_ALLREDUCE=nccl
cat $O/r2_call4.rc; tail -4 $O/r2_tests4.log
