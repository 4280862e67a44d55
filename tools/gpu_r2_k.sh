#!/bin/bash
# round 2, build k: full GPU suite, X1-variant bit-exact check, per-op rates, default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
SPF_B200_LIB=variants/libspf_x1on.so timeout 600 python -m pytest tests/test_gpu_bitexact.py -x -q -m gpu -k "pbs_pair or cbs" 2>&1 | tail -2
timeout 900 python tools/op_bench.py 4096 > gpurun_out/r2_k_op_bench.jsonl 2> gpurun_out/k_op.err; echo "op_bench rc=$?"
grep -h "rlwe\|cpu_port" gpurun_out/r2_k_op_bench.jsonl | cut -c1-400
timeout 1200 python bench.py > gpurun_out/r2_k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/r2_k_bench.json
