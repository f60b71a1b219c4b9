#!/bin/bash
# round 2, GPU call 7 (1 GPU): tcgen05 attention forward - kernel tests, full suite, bench A/B (pretrain + retrieval)
O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" > $O/r2_tests7_attn.log 2>&1; echo "attn tests rc=$?" > $O/r2_call7.rc
timeout 900 python -m pytest tests -m gpu -q > $O/r2_tests7.log 2>&1; echo "tests rc=$?" >> $O/r2_call7.rc
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-optimizer-leg > $O/r2_b7_tc.json 2> $O/r2_b7_tc.err; echo "bench tc rc=$?" >> $O/r2_call7.rc
VOLTA_B200_ATTN_TC=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg > $O/r2_b7_notc.json 2> $O/r2_b7_notc.err; echo "bench notc rc=$?" >> $O/r2_call7.rc
timeout 300 python bench.py --workload retrieval --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_b7_retr_tc.json 2> $O/r2_b7_retr_tc.err; echo "retr tc rc=$?" >> $O/r2_call7.rc
VOLTA_B200_ATTN_TC=0 timeout 300 python bench.py --workload retrieval --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2_b7_retr_notc.json 2> $O/r2_b7_retr_notc.err; echo "retr notc rc=$?" >> $O/r2_call7.rc
cat $O/r2_call7.rc; tail -5 $O/r2_tests7_attn.log; tail -3 $O/r2_tests7.log
