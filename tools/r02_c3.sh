# c3 lane drivers and lane counts (staged builds by default on throughput-mode handles)
for D in native python; do
 for L in 8 16; do
  echo "== c3 driver $D lanes $L"
  python bench.py --workload c3 --steps 4096 --no-cpu-baseline --c3-lanes $L --c3-driver $D 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'launches/pair',d['gpu_launches']/d['steps'],'max_err',d['max_translation_error_vs_truth_m'])"
 done
done
python -m pytest tests/test_gpu_bench.py -m gpu -x -q 2>&1 | tail -3
