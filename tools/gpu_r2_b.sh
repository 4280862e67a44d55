#!/bin/bash
# GPU call B: BSK ring + fused-store variants of pbs_kernel: bit-exactness vs the emulator, then timing
mkdir -p gpurun_out
for v in ring3c ring3cf fs3c; do
  SPF_B200_LIB=$PWD/variants/libspf_$v.so timeout 300 python -m pytest tests/test_gpu_bitexact.py -q -x -k "pair or cbs" > gpurun_out/b_bitexact_$v.log 2>&1; echo "$v bitexact rc=$?" | tee -a gpurun_out/b_summary.log
done
for v in tr3c fs3 fs3c ring3 ring3c ring3cf; do
  echo "== variant $v" >> gpurun_out/b_sweep.log
  SPF_B200_LIB=$PWD/variants/libspf_$v.so tools/wave_sweep.sh "4096 444" >> gpurun_out/b_sweep.log 2>&1
done
cat gpurun_out/b_sweep.log; tail -3 gpurun_out/b_bitexact_*.log
