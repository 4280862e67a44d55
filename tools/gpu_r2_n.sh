#!/bin/bash
# round 2, build n (final): full GPU suite, default bench, ncu launch list of the same command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 1500 python bench.py > gpurun_out/r2_n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_n_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value']), 'graphs', d['e2e']['graphs_per_step'],
      'frac', round(d['roofline']['frac'], 4), 'check', d['check'], 'add32', d['parasol_add_latency']['add32']['gpu_ms'], 'mul32', d['parasol_mul32_cmp_latency']['gpu_ms'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_n_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-add --no-sweep --no-e2e --check 0 > gpurun_out/n_ncu_list.log 2>&1; echo "launch list rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_n_reference.json 2> gpurun_out/n_ref.err; echo "reference rc=$?"; cut -c1-300 gpurun_out/r2_n_reference.json
