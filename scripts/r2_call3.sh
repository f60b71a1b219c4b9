#!/bin/bash
# round 2, GPU call 3 (2 GPUs): multi-rank DDP numerical check (fp32 / bf16 wire), then 2-GPU bench A/B of the wire format and
# of the NCCL buffer registration
O=gpurun_out
python -m pytest tests/test_parallel_gpu.py -m gpu -q -s > $O/r2_ddp_tests.log 2>&1; echo "ddp tests rc=$?" > $O/r2_call3.rc
run() {  # name, env...
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-optimizer-leg --no-full-length-leg > $O/r2_b3_$name.json 2> $O/r2_b3_$name.err; echo "$name rc=$?" >> $O/r2_call3.rc
}
run reg_fp32 VOLTA_B200_DDP_REGISTER=1
run noreg_fp32 VOLTA_B200_DDP_REGISTER=0
run reg_bf16 VOLTA_B200_DDP_REGISTER=1 VOLTA_B200_DDP_BF16=1
run noreg_bf16 VOLTA_B200_DDP_REGISTER=0 VOLTA_B200_DDP_BF16=1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-optimizer-leg --no-full-length-leg > $O/r2_b3_1gpu.json 2> $O/r2_b3_1gpu.err; echo "1gpu rc=$?" >> $O/r2_call3.rc
cat $O/r2_call3.rc; tail -4 $O/r2_ddp_tests.log
