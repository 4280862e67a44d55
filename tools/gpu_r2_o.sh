#!/bin/bash
mkdir -p gpurun_out
for flags in "--steps 3 --warmup 3 --no-cpu-baseline --no-add --no-sweep --check 64" "--steps 5 --warmup 3 --no-cpu-baseline --no-add --no-sweep --check 64" "--steps 3 --warmup 3 --no-cpu-baseline --no-add --check 64" "--steps 3 --warmup 3 --no-add --no-sweep --check 64" "--steps 3 --warmup 3 --no-cpu-baseline --no-sweep --check 64" ""; do
  for ov in 1 0; do
  SPF_B200_CBS_OVERLAP=$ov timeout 900 python bench.py $flags 2> gpurun_out/o_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ov=$ov flags=[$flags]', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'sweep', [round(r['cbs_per_s']) for r in (d.get('throughput_sweep') or [])])"
  done
done | tee gpurun_out/r2_o_flags.txt
