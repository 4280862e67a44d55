#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/f_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/f_bench.json; tail -5 gpurun_out/f_bench.err
