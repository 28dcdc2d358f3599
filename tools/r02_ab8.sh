python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sharded 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'results_only',round(d['e2e']['results_only']['value']),'single_call',round(d['e2e']['single_call']['value']),'latency_ms',d['latency']['ms_per_align'])"
python bench.py --workload c1 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print({k:(v['single_ms'],v['ktimes_ms'],v['fitness']) for k,v in d['methods'].items()})"
