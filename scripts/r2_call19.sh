#!/bin/bash
# round 2, call 19: final build - GPU suite, smoke, default bench line
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/c19_tests.log 2>&1; echo "tests rc=$?" >> $O/c19_tests.log
tail -3 $O/c19_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/c19_smoke.log 2>&1; echo "smoke rc=$?" >> $O/c19_smoke.log; tail -2 $O/c19_smoke.log
timeout 600 python bench.py > $O/c19_bench.json 2> $O/c19_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$O/c19_bench.json'));print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['gpu_launches'])"
timeout 300 python bench.py --batch 32 --seq-len 38 --no-cpu-baseline --no-optimizer-leg --no-full-length-leg > $O/c19_bench_b32.json 2> $O/c19_bench_b32.err; echo "bench b32 rc=$?"
python -c "import json;d=json.load(open('$O/c19_bench_b32.json'));print(d['value'], d['ms_per_step'], d['e2e'])"
