#!/bin/bash
O=gpurun_out
for be in tc mma; do
  python scripts/attn_one.py $be 256 36 37 12 64 9 packed >> $O/r2_attn_times.log 2>&1
  python scripts/attn_one.py $be 256 36 37 12 64 6 packed >> $O/r2_attn_times.log 2>&1
  python scripts/attn_one.py $be 256 36 37 12 64 1 packed >> $O/r2_attn_times.log 2>&1
  python scripts/attn_one.py $be 256 36 37 12 64 15 packed >> $O/r2_attn_times.log 2>&1
  python scripts/attn_one.py $be 500 30 101 8 128 8 >> $O/r2_attn_times.log 2>&1
  python scripts/attn_one.py $be 500 30 101 8 128 6 >> $O/r2_attn_times.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -c 2 -o $O/r2_attn_tc_pre python scripts/attn_one.py tc 256 36 37 12 64 9 packed > $O/r2_ncu_attn1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -c 1 -o $O/r2_attn_tc_retr python scripts/attn_one.py tc 500 30 101 8 128 8 > $O/r2_ncu_attn2.log 2>&1
cat $O/r2_attn_times.log
