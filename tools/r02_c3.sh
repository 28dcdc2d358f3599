for B in staged fused; do
 for L in 8 16; do
  echo "== c3 build path $B lanes $L"
  NDTB200_BUILD_PATH=$B python bench.py --workload c3 --steps 4096 --no-cpu-baseline --c3-lanes $L 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value',d['value'],'ms_per_step',d['ms_per_step'],'launches/pair',d['gpu_launches']/d['steps'])"
 done
done
echo "== mapper"
python bench.py --workload mapper --steps 200 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value',d['value'],'ms_per_step',d['ms_per_step'], d['step_ms_percentiles'])"
