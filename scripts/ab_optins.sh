#!/usr/bin/env bash
# A/B of the switches that were built and emulation-checked but not yet run on a B200 (DESIGN.md 6a / 8 "Next").
# One GPU:   bash scripts/ab_optins.sh single      (about 4 minutes)
# N GPUs:    bash scripts/ab_optins.sh ddp N        (run under `gpurun --gpus N`; about 2 minutes)
# Everything lands in gpurun_out/ab_*.{log,json}; numbers printed under a profiler are never used.
set -u
mkdir -p gpurun_out
mode=${1:-single}
B="--steps 20 --warmup 5 --no-cpu-baseline --no-optimizer-leg"
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "value %.0f samples/s  %.3f ms/step  e2e %s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value")))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
if [ "$mode" = single ]; then
  # 1. packed rows for shared sub-layers / any B*T: parity first, then throughput per family
  VOLTA_B200_PACK_SHARED=1 VOLTA_B200_PACK_ANY_ROWS=1 python -m pytest tests/test_model_gpu.py -m gpu -q -k "shared_sublayers or packed_rows" \
      > gpurun_out/ab_pack_tests.log 2>&1; tail -3 gpurun_out/ab_pack_tests.log
  for model in ctrl_lxmert ctrl_visualbert_base; do
    for on in 0 1; do
      VOLTA_B200_PACK_SHARED=$on python bench.py $B --no-e2e --model $model > gpurun_out/ab_${model}_pack$on.json 2> gpurun_out/ab_${model}_pack$on.err
      line gpurun_out/ab_${model}_pack$on.json
    done
  done
  # 2. the reference recipe's per-GPU batch (32 x 36 rows: not a multiple of 256)
  for on in 0 1; do
    VOLTA_B200_PACK_ANY_ROWS=$on python bench.py $B --no-e2e --batch 32 > gpurun_out/ab_b32_anyrows$on.json 2> gpurun_out/ab_b32_anyrows$on.err
    line gpurun_out/ab_b32_anyrows$on.json
  done
  # 3. end to end from the raw detector output, assembled on the device
  for raw in fp32 bf16; do
    python bench.py $B --raw-input $raw > gpurun_out/ab_raw_$raw.json 2> gpurun_out/ab_raw_$raw.err
    line gpurun_out/ab_raw_$raw.json
  done
else
  N=${2:-2}
  for on in 0 1; do
    VOLTA_B200_DDP_BF16=$on python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N $B --no-e2e > gpurun_out/ab_ddp${N}_bf16wire$on.json 2> gpurun_out/ab_ddp${N}_bf16wire$on.err
    line gpurun_out/ab_ddp${N}_bf16wire$on.json
  done
fi
