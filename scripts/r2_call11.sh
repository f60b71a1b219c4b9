#!/bin/bash
O=gpurun_out
VOLTA_B200_BENCH_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg > $O/r2_b11_pre.json 2> $O/r2_b11_pre_verbose.err; echo "pre rc=$?" > $O/r2_call11.rc
timeout 300 python bench.py --model ctrl_visualbert_base --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg > $O/r2_b11_vb.json 2> $O/r2_b11_vb.err; echo "vb rc=$?" >> $O/r2_call11.rc
timeout 300 python bench.py --model ctrl_lxmert --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg > $O/r2_b11_lx.json 2> $O/r2_b11_lx.err; echo "lx rc=$?" >> $O/r2_call11.rc
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_default_plan_gpu.py -m gpu -q -x > $O/r2_tests11.log 2>&1; echo "tests rc=$?" >> $O/r2_call11.rc
cat $O/r2_call11.rc; tail -3 $O/r2_tests11.log; grep "^gemm" $O/r2_b11_pre_verbose.err | head -30
