#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-add --no-sweep --check 16 > gpurun_out/i_bench_plain.json 2> gpurun_out/i_bench_plain.err; echo "plain bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_i_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-add --no-sweep --no-e2e --check 0 > gpurun_out/i_ncu_list.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pbs_kernel -c 1 -f -o gpurun_out/r2_i_pbs_b4096 python tools/cbs_time.py 4096 1 > gpurun_out/i_ncu_pbs.log 2>&1; echo "pbs ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pbs_kernel -c 1 -f -o gpurun_out/r2_i_pbs_b444 python tools/cbs_time.py 444 1 > gpurun_out/i_ncu_pbs444.log 2>&1; echo "pbs444 ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_ss -c 1 -f -o gpurun_out/r2_i_trace_b4096 python tools/cbs_time.py 4096 1 > gpurun_out/i_ncu_trace.log 2>&1; echo "trace ncu rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_i_launches.csv
