for w in 1 2 3 5; do
SPF_B200_CHUNK_WAVES=$w python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-add --check 2 > gpurun_out/e2e_$w.log 2>/dev/null
python -c "
import json;d=json.loads(open('gpurun_out/e2e_$w.log').read().strip().splitlines()[-1]);print('waves',$w,'device',round(d['value']),'e2e',round(d['e2e']['value']))"
done
