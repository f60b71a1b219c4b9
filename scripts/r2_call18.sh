#!/bin/bash
# round 2, call 18: task criteria + RAdam on the GPU
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_task_utils.py tests/test_radam.py -m gpu -q > $O/c18_tests.log 2>&1; echo "rc=$?" | tee -a $O/c18_tests.log
grep -v "^model's" $O/c18_tests.log | tail -40
