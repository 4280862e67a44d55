#!/bin/bash
# final robustness pass: GPU suite three times, smoke, default bench twice
mkdir -p gpurun_out
for i in 1 2 3; do timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -1; done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for i in 1 2; do
timeout 1500 python bench.py > gpurun_out/r2_u_bench_$i.json 2> gpurun_out/u_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/r2_u_bench_$i.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'], 4), 'ok', d['check']['ok'], d['check']['e2e_ok'],
      'add32', round(d['parasol_add_latency']['add32']['gpu_ms'], 3), 'add32_bdd', round(d['parasol_add_latency']['add32_bdd_circuit']['gpu_ms'], 3), 'mul32', round(d['parasol_mul32_cmp_latency']['gpu_ms'], 2), 'cpu', round(d['cpu_baseline']['value']), 'sweep', [round(r['cbs_per_s']) for r in d['throughput_sweep']])
PY
done
