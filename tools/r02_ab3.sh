python -m pytest tests -m gpu -x -q -k "batch or throughput" 2>&1 | tail -3
for M in divide share; do
  echo "== batch mode $M"
  NDTB200_BATCH_MODE=$M python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sharded 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value',d['value'],'ms_per_step',d['ms_per_step'],'lat_ms',d['latency']['ms_per_align'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])"
done
for R in 16 32 128; do
  echo "== replicas $R (divide)"
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sharded --replicas $R 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value',d['value'],'ms_per_step',d['ms_per_step'],'e2e',d['e2e']['value'])"
done
