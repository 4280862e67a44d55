#!/bin/bash
# usage: tools/run_modes.sh NGPUS "32 4 3 mul" ...   -- runs each workload with the nccl and the peer exchange
n=$1; shift
for k in "$@"; do
  for x in nccl peer; do
    timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 tools/sharded_graph_run.py $k $x 2>&1 | tail -1 |
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['workload'], '| gpus', d['n_gpus'], d['exchange'], '| ms', round(d['graph_ms_max_over_ranks'],2), '| correct', d['correct_on_all_ranks'], d['outputs_checked_over_ranks'], '/', d['outputs'])"
  done
done
