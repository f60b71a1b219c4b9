#!/bin/bash
# round 2, GPU call 1: GPU tests with the packed plan enabled for shared sub-layers / any row count, then the bench legs
export VOLTA_B200_PACK_SHARED=1 VOLTA_B200_PACK_ANY_ROWS=1
O=gpurun_out
python -m pytest tests -m gpu -q -s > $O/r2_tests1.log 2>&1; echo "tests rc=$?" > $O/r2_call1.rc
python bench.py --steps 20 --warmup 5 > $O/r2_b1_vilbert.json 2> $O/r2_b1_vilbert.err; echo "vilbert rc=$?" >> $O/r2_call1.rc
python bench.py --model ctrl_visualbert_base --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_b1_visualbert.json 2> $O/r2_b1_visualbert.err; echo "visualbert rc=$?" >> $O/r2_call1.rc
VOLTA_B200_PACK_SHARED=0 python bench.py --model ctrl_visualbert_base --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg > $O/r2_b1_visualbert_nopack.json 2> $O/r2_b1_visualbert_nopack.err; echo "visualbert nopack rc=$?" >> $O/r2_call1.rc
python bench.py --model ctrl_lxmert --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_b1_lxmert.json 2> $O/r2_b1_lxmert.err; echo "lxmert rc=$?" >> $O/r2_call1.rc
VOLTA_B200_PACK_SHARED=0 python bench.py --model ctrl_lxmert --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg > $O/r2_b1_lxmert_nopack.json 2> $O/r2_b1_lxmert_nopack.err; echo "lxmert nopack rc=$?" >> $O/r2_call1.rc
python bench.py --batch 32 --seq-len 38 --steps 30 --warmup 5 --no-cpu-baseline > $O/r2_b1_recipe32.json 2> $O/r2_b1_recipe32.err; echo "recipe32 rc=$?" >> $O/r2_call1.rc
VOLTA_B200_PACK_ANY_ROWS=0 python bench.py --batch 32 --seq-len 38 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg > $O/r2_b1_recipe32_nopack.json 2> $O/r2_b1_recipe32_nopack.err; echo "recipe32 nopack rc=$?" >> $O/r2_call1.rc
python bench.py --workload retrieval --steps 10 --warmup 3 > $O/r2_b1_retrieval.json 2> $O/r2_b1_retrieval.err; echo "retrieval rc=$?" >> $O/r2_call1.rc
cat $O/r2_call1.rc; tail -5 $O/r2_tests1.log
