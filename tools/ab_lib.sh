#!/bin/bash
# A/B of two builds of the library on the same box: tools/ab_lib.sh OTHER.so "batches"  (alternates 3 times)
for rep in 1 2 3; do
  for lib in "" "$1"; do
    SPF_B200_LIB=$lib timeout 200 python tools/throughput_sweep.py --batches $2 2>&1 | grep batch_per |
      python -c "import sys,json; print('${lib:-default}', ' '.join('%d:%.3f' % (json.loads(l)['batch_per_gpu'], json.loads(l)['pbs_ms']) for l in sys.stdin))"
  done
done
