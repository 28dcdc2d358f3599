for C in 8 32; do
 for M in divide share; do
  echo "== connections $C batch mode $M"
  CUDA_DEVICE_MAX_CONNECTIONS=$C NDTB200_BATCH_MODE=$M python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sharded 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value',d['value'],'ms_per_step',d['ms_per_step'],'lat_ms',d['latency']['ms_per_align'],'e2e',d['e2e']['value'],'single',d['e2e']['single_call']['value'])"
 done
done
echo "== connections 32 divide sorted"
CUDA_DEVICE_MAX_CONNECTIONS=32 NDTB200_SORT_SOURCE_MIN=1000 NDTB200_BATCH_MODE=divide python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sharded 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value',d['value'],'ms_per_step',d['ms_per_step'],'lat_ms',d['latency']['ms_per_align'],'e2e',d['e2e']['value'],'single',d['e2e']['single_call']['value'])"
