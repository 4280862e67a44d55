#!/bin/bash
mkdir -p gpurun_out
for n in 0 1 2 3 4 8 16 32 64 127 124 123 119 111 95 63; do
  SPF_B200_LIB=$PWD/variants/libspf_abl$n.so timeout 120 python tools/pbs_time.py 444 4 2>&1 | tail -1 >> gpurun_out/c_ablate.log
done
cat gpurun_out/c_ablate.log
