#!/bin/bash
# round 2, call 25: last build - smoke + the driver-style default bench line
O=gpurun_out; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/c25_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c25_smoke.log
timeout 600 python bench.py > $O/c25_bench.json 2> $O/c25_bench.err; echo "bench rc=$? lines=$(wc -l < $O/c25_bench.json)"
python -c "import json;d=json.loads(open('$O/c25_bench.json').read());print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'], d['clocks'])"
