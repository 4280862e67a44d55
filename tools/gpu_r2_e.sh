#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/e_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/e_bench.json; tail -3 gpurun_out/e_bench.err
SPF_B200_PBS_TAIL=pair timeout 300 python tools/pbs_time.py 444,4096 4
timeout 300 python tools/pbs_time.py 444,4096 4
