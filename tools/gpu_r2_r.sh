#!/bin/bash
# round 2, build r: runs of MUX-tree levels in one cooperative launch (cmux_chain_kernel)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_graph.py tests/test_gpu_async.py -x -q -m gpu 2>&1 | tail -3
for rep in 1 2 3; do
  for nc in 0 1; do
    echo "NO_CHAIN=$nc"
    SPF_B200_NO_CHAIN=$nc timeout 300 python examples/mul_cmp.py 2>&1 | tail -1
  done
done | tee gpurun_out/r2_r_chain_ab.txt
