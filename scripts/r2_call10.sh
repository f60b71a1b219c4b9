#!/bin/bash
O=gpurun_out
timeout 300 python bench.py --workload retrieval --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_b10_retr.json 2> $O/r2_b10_retr.err; echo "retr rc=$?" > $O/r2_call10.rc
VOLTA_B200_BENCH_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg > $O/r2_b10_pre.json 2> $O/r2_b10_pre_verbose.err; echo "pre rc=$?" >> $O/r2_call10.rc
timeout 600 python -m pytest tests/test_retrieval.py tests/test_pipeline.py -m gpu -q > $O/r2_tests10.log 2>&1; echo "tests rc=$?" >> $O/r2_call10.rc
cat $O/r2_call10.rc; tail -3 $O/r2_tests10.log; grep "^gemm" $O/r2_b10_pre_verbose.err | head -40
