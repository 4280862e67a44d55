#!/bin/bash
mkdir -p gpurun_out
for v in tr3c tr3ct2 ring3c ring3ct2; do
  SPF_B200_LIB=$PWD/variants/libspf_$v.so timeout 200 python tools/pbs_time.py 444,4096 4 2>&1 | tail -1 >> gpurun_out/d_time.log
done
cat gpurun_out/d_time.log
SPF_B200_LIB=$PWD/variants/libspf_ring3c.so timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest_ring3c.log 2>&1; echo "ring3c pytest rc=$?"; tail -3 gpurun_out/d_pytest_ring3c.log
