#!/bin/bash
# GPU call A of round 2: parity (incl. GPU-vs-emulator bit-exact), variant sweep of pbs_kernel, ncu of the 4-pair build
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
SPF_B200_LIB=$PWD/variants/libspf_tr4.so timeout 300 python -m pytest tests/test_gpu_bitexact.py -q -k "pair or cbs" > gpurun_out/a_pytest_tr4.log 2>&1; echo "rc=$?" >> gpurun_out/a_pytest_tr4.log
for v in default base3 tr4 tr3c; do
  echo "== variant $v" >> gpurun_out/a_sweep.log
  if [ $v = default ]; then lib=""; else lib=$PWD/variants/libspf_$v.so; fi
  SPF_B200_LIB=$lib tools/wave_sweep.sh "4096 592 444" >> gpurun_out/a_sweep.log 2>&1
done
SPF_B200_LIB=$PWD/variants/libspf_tr4.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:pbs_kernel -c 1 -f -o gpurun_out/r2a_tr4_pbs python tools/throughput_sweep.py --batches 592 > gpurun_out/a_ncu_tr4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pbs_kernel -c 1 -f -o gpurun_out/r2a_tr3_pbs python tools/throughput_sweep.py --batches 444 > gpurun_out/a_ncu_tr3.log 2>&1
ls -la gpurun_out/ | tail -20
tail -5 gpurun_out/a_pytest.log; cat gpurun_out/a_pytest_tr4.log | tail -3; cat gpurun_out/a_sweep.log
