#!/bin/bash
# last pass of round 2 (after the instruction-count work): suite, smoke, default bench, driver-style bench, reference arm, launch list, ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/r2_zz_bench.json 2> gpurun_out/zz_bench.err; echo "bench rc=$?"
timeout 1500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-add > gpurun_out/r2_zz_bench_steps20.json 2>> gpurun_out/zz_bench.err; echo "bench20 rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_zz_reference.json 2> gpurun_out/zz_ref.err; echo "reference rc=$?"
python - <<PY
import json
for f in ('gpurun_out/r2_zz_bench.json', 'gpurun_out/r2_zz_bench_steps20.json'):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'], 4), 'ok', d['check']['ok'], d['check']['e2e_ok'],
          'sweep', [round(r['cbs_per_s']) for r in d['throughput_sweep']])
    if d.get('parasol_add_latency'):
        print('  add32', round(d['parasol_add_latency']['add32']['gpu_ms'], 3), 'add32_bdd', round(d['parasol_add_latency']['add32_bdd_circuit']['gpu_ms'], 3), 'mul32', round(d['parasol_mul32_cmp_latency']['gpu_ms'], 2), 'cpu', round(d['cpu_baseline']['value']))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_zz_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-add --no-sweep --no-e2e --check 0 > gpurun_out/zz_ncu_list.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pbs_kernel -c 1 -f -o gpurun_out/r2_zz_pbs_b444 python tools/cbs_time.py 444 1 > gpurun_out/zz_ncu_pbs444.log 2>&1; echo "pbs444 ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pbs_kernel -c 1 -f -o gpurun_out/r2_zz_pbs_b4096 python tools/cbs_time.py 4096 1 > gpurun_out/zz_ncu_pbs4096.log 2>&1; echo "pbs4096 ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_ss -c 1 -f -o gpurun_out/r2_zz_trace_b4096 python tools/cbs_time.py 4096 1 > gpurun_out/zz_ncu_trace.log 2>&1; echo "trace ncu rc=$?"
ls -la gpurun_out/r2_zz_*
