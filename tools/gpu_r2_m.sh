#!/bin/bash
# round 2, build m: cmux_wide with the selector GGSW staged in tensor memory during the forward transforms
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_bitexact.py tests/test_gpu_graph.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for rep in 1 2 3; do
  for lib in "" variants/libspf_nostage.so; do
    echo "lib=${lib:-default}"
    SPF_B200_LIB=$lib timeout 300 python examples/mul_cmp.py 2>&1 | tail -1
    SPF_B200_LIB=$lib timeout 300 python tools/add_latency.py 32 4 2>&1 | tail -3 | head -2
  done
done | tee gpurun_out/r2_m_wide_stage_ab.txt
