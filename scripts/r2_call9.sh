#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2_tests9.log 2>&1; echo "tests rc=$?" > $O/r2_call9.rc
rm -f $O/r2_attn_times2.log
for be in tc mma; do
  python scripts/attn_one.py $be 500 30 101 8 128 8 >> $O/r2_attn_times2.log 2>&1
  python scripts/attn_one.py $be 500 30 101 8 128 6 >> $O/r2_attn_times2.log 2>&1
  python scripts/attn_one.py $be 500 30 101 8 128 1 >> $O/r2_attn_times2.log 2>&1
  python scripts/attn_one.py $be 256 36 37 12 64 9 packed >> $O/r2_attn_times2.log 2>&1
done
timeout 300 python bench.py --workload retrieval --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_b9_retr.json 2> $O/r2_b9_retr.err; echo "retr rc=$?" >> $O/r2_call9.rc
VOLTA_B200_ATTN_TC=0 timeout 300 python bench.py --workload retrieval --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2_b9_retr_notc.json 2> $O/r2_b9_retr_notc.err; echo "retr notc rc=$?" >> $O/r2_call9.rc
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_b9_pre.json 2> $O/r2_b9_pre.err; echo "pre rc=$?" >> $O/r2_call9.rc
cat $O/r2_call9.rc; tail -3 $O/r2_tests9.log; cat $O/r2_attn_times2.log
