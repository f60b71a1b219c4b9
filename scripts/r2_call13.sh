#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > $O/r2_tests13.log 2>&1; echo "tests rc=$?" > $O/r2_call13.rc
cat $O/r2_call13.rc; tail -3 $O/r2_tests13.log; grep GRADTOL $O/r2_tests13.log
