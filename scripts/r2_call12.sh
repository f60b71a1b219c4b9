#!/bin/bash
# round 2, final 1-GPU evidence run: tests, smoke, bench lines of every BASELINE config that fits one GPU, reference arm,
# ncu launch list (+DRAM bytes) of one step, per-kernel timings and one ncu --set full pass over the hot kernels
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > $O/r2_tests12.log 2>&1; echo "tests rc=$?" > $O/r2_call12.rc
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke12.log 2>&1; echo "smoke rc=$?" >> $O/r2_call12.rc
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > $O/r2_b12_reference.json 2> $O/r2_b12_reference.err; echo "ref rc=$?" >> $O/r2_call12.rc
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2_b12_vilbert.json 2> $O/r2_b12_vilbert.err; echo "vilbert rc=$?" >> $O/r2_call12.rc
timeout 300 python bench.py --model ctrl_visualbert_base --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_b12_visualbert.json 2> $O/r2_b12_visualbert.err; echo "visualbert rc=$?" >> $O/r2_call12.rc
timeout 300 python bench.py --model ctrl_lxmert --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_b12_lxmert.json 2> $O/r2_b12_lxmert.err; echo "lxmert rc=$?" >> $O/r2_call12.rc
timeout 300 python bench.py --batch 32 --seq-len 38 --steps 30 --warmup 5 --no-cpu-baseline > $O/r2_b12_recipe32.json 2> $O/r2_b12_recipe32.err; echo "recipe32 rc=$?" >> $O/r2_call12.rc
timeout 300 python bench.py --workload retrieval --steps 10 --warmup 3 > $O/r2_b12_retrieval.json 2> $O/r2_b12_retrieval.err; echo "retrieval rc=$?" >> $O/r2_call12.rc
timeout 300 python bench.py --full-length --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-optimizer-leg > $O/r2_b12_full_length.json 2> $O/r2_b12_full_length.err; echo "full-length rc=$?" >> $O/r2_call12.rc
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv --log-file $O/r2_launches_bench_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg > $O/r2_ncu_step.log 2>&1; echo "ncu step rc=$?" >> $O/r2_call12.rc
timeout 300 python scripts/opbench.py > $O/r2_opbench_per_kernel.txt 2>&1; echo "opbench rc=$?" >> $O/r2_call12.rc
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c 26 -o $O/r2_ops_full python scripts/opbench.py --once > $O/r2_ncu_ops.log 2>&1; echo "ncu ops rc=$?" >> $O/r2_call12.rc
cat $O/r2_call12.rc; tail -3 $O/r2_tests12.log; tail -2 $O/r2_smoke12.log
