#!/bin/bash
# round 2, build l: CBS with the partial last wave of the blind rotation overlapped with the trace kernels (launch_cbs)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for ov in 1 0 1 0; do
  SPF_B200_CBS_OVERLAP=$ov timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-add --no-sweep --check 64 2> gpurun_out/l_bench_$ov.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('overlap=$ov', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'check', d['check']['ok'], d['check'].get('e2e_ok'))"
done | tee gpurun_out/r2_l_overlap_ab.txt
