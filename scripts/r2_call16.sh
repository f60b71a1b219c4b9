#!/bin/bash
# round 2, call 16: GPU suite with the attention-data tests + smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c16_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c16_tests.log
tail -5 gpurun_out/c16_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c16_smoke.log 2>&1; tail -2 gpurun_out/c16_smoke.log
