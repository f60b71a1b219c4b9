#!/bin/bash
# round 2, GPU call 2: GPU tests with the new defaults, launch lists (ncu) of the headline step and of the retrieval step
O=gpurun_out
python -m pytest tests -m gpu -q -s > $O/r2_tests2.log 2>&1; echo "tests rc=$?" > $O/r2_call2.rc
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke2.log 2>&1; echo "smoke rc=$?" >> $O/r2_call2.rc
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg > $O/r2_b2_pre.json 2> $O/r2_b2_pre.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2_launches_vilbert.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg > $O/r2_ncu_vilbert.log 2>&1; echo "ncu vilbert rc=$?" >> $O/r2_call2.rc
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_launches_retrieval.csv python bench.py --workload retrieval --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2_ncu_retrieval.log 2>&1; echo "ncu retrieval rc=$?" >> $O/r2_call2.rc
cat $O/r2_call2.rc; tail -5 $O/r2_tests2.log
