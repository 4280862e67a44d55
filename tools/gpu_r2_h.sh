#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bitexact.py tests/test_gpu_parity.py -q -x 2>&1 | tail -3
SPF_B200_LIB=$PWD/variants/libspf_trnotw.so timeout 200 python tools/cbs_time.py 4096 4 | tee -a gpurun_out/h_time.log
timeout 200 python tools/cbs_time.py 4096 4 | tee -a gpurun_out/h_time.log
