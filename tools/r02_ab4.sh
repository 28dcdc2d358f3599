for M in divide share; do
 for S in 1000 262144; do
  echo "== batch mode $M sort_min $S"
  NDTB200_SORT_SOURCE_MIN=$S NDTB200_BATCH_MODE=$M python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sharded 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value',d['value'],'ms_per_step',d['ms_per_step'],'lat_ms',d['latency']['ms_per_align'],'e2e',d['e2e']['value'],'single',d['e2e']['single_call']['value'])"
 done
done
