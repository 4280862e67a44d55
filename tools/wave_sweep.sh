#!/bin/bash
# usage: tools/wave_sweep.sh "148 296 444" -- prints pbs_kernel ms per launch for each batch size
for b in $1; do
  python bench.py --batch $b --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-add --check 4 > gpurun_out/q$b.log 2>gpurun_out/q$b.err
  python -c "
import json;d=json.loads(open('gpurun_out/q$b.log').read().strip().splitlines()[-1]);r=d['roofline'];print('batch',$b,'pbs_ms',round(r['ms_per_launch'],3),'frac',round(r['frac'],4),'step_ms',round(d['ms_per_step'],3),'cbs/s',round(d['value']),d['check'])"
done
