#!/bin/bash
# round 2, call 24: aux chunks two ahead in the GEMM epilogue - kernel tests, A/B against the previous library, full suite
O=gpurun_out; mkdir -p $O
L=volta_b200/libvolta_b200.so
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" > $O/c24_gemm_tests.log 2>&1; echo "gemm tests rc=$?"; tail -2 $O/c24_gemm_tests.log
F="--no-cpu-baseline --no-e2e --no-optimizer-leg --no-full-length-leg --steps 30 --warmup 5"
cp $L /tmp/new.so
for v in new old new old; do
  if [ $v = old ]; then cp volta_b200/libvolta_b200_prev.so $L; else cp /tmp/new.so $L; fi
  VOLTA_B200_BENCH_VERBOSE=1 timeout 300 python bench.py $F > $O/c24_$v.json 2> $O/c24_$v.err; echo "$v rc=$?"
  python -c "import json;d=json.loads(open('$O/c24_$v.json').read());print('$v', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])"
  grep -E "epi5|K768 a1 b1 epi4|K768 a1 b0 epi0" $O/c24_$v.err | head -6
done
cp /tmp/new.so $L
timeout 900 python -m pytest tests -m gpu -x -q > $O/c24_tests.log 2>&1; echo "tests rc=$?" >> $O/c24_tests.log; tail -3 $O/c24_tests.log
