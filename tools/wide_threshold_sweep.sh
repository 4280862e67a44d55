#!/bin/bash
# In-graph A/B of the wide-CMUX threshold (SPF_B200_CMUX_WIDE_MAX) on the mul32 programs, one GPU.
# usage: tools/wide_threshold_sweep.sh "296 592 1184" "4 8"
for m in $1; do
  for P in $2; do
    SPF_B200_CMUX_WIDE_MAX=$m timeout 200 python tools/sharded_graph_run.py 32 $P 3 mul 2>&1 | tail -1 |
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('wide_max', $m, 'programs', $P, 'ms', round(d['graph_ms_max_over_ranks'],2), d['correct_on_all_ranks'])"
  done
done
