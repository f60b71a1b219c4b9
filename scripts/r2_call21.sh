#!/bin/bash
# round 2, call 21: bench.py after the stdout change - driver-style default run and the reference arm
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py > $O/c21_bench.json 2> $O/c21_bench.err; echo "bench rc=$? lines=$(wc -l < $O/c21_bench.json)"
python -c "import json;d=json.loads(open('$O/c21_bench.json').read());print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/c21_ref.json 2> $O/c21_ref.err; echo "ref rc=$? lines=$(wc -l < $O/c21_ref.json)"
python -c "import json;d=json.loads(open('$O/c21_ref.json').read());print(d['impl'], d['value'], d['cpu_baseline']['cores'])"
