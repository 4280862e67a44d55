#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for i in 1 2 3; do
timeout 1500 python bench.py --no-cpu-baseline --no-add 2> gpurun_out/p_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('run $i value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'sweep', [round(r['cbs_per_s']) for r in (d.get('throughput_sweep') or [])], d['check']['ok'], d['check']['e2e_ok'])"
done
