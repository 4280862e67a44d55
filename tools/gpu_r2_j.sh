#!/bin/bash
# round 2, build j: first exchange through tensor memory (SPF_PBS_TMEM_X1) -- bit-exact tests, then same-box A/B against the build without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bitexact.py -x -q -m gpu 2>&1 | tail -5
for rep in 1 2 3; do
  for lib in "" variants/libspf_x1off.so; do
    SPF_B200_LIB=$lib timeout 300 python tools/pbs_time.py 444,4096 5 2>&1 | tail -1
  done
done | tee gpurun_out/r2_j_x1_ab.txt
