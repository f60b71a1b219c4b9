#!/bin/bash
# round 2, call 17: FusedRAdam on the GPU; weight-gradient GEMM on the side stream (A/B, 3 variants, two batch sizes)
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_radam.py -m gpu -x -q > $O/c17_radam.log 2>&1; echo "radam rc=$?" | tee -a $O/c17_radam.log
F="--no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg --steps 30 --warmup 5"
for w in 0 1 2 0; do
  VOLTA_B200_WGRAD_SIDE=$w timeout 300 python bench.py $F > $O/c17_b256_w$w.json 2> $O/c17_b256_w$w.err; echo "b256 w$w rc=$?"
  python -c "import json;d=json.load(open('$O/c17_b256_w$w.json'));print('B256 side=$w', d['value'], d['ms_per_step'])"
done
for w in 0 1 2; do
  VOLTA_B200_WGRAD_SIDE=$w timeout 300 python bench.py $F --batch 32 --seq-len 38 > $O/c17_b32_w$w.json 2> $O/c17_b32_w$w.err; echo "b32 w$w rc=$?"
  python -c "import json;d=json.load(open('$O/c17_b32_w$w.json'));print('B32 side=$w', d['value'], d['ms_per_step'])"
done
for w in 1 2; do
  VOLTA_B200_WGRAD_SIDE=$w timeout 600 python -m pytest tests/test_default_plan_gpu.py -x -q > $O/c17_plan_w$w.log 2>&1; echo "plan w$w rc=$?"; tail -1 $O/c17_plan_w$w.log
done
