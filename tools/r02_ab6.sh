# throughput-shape CTA size A/B: 256 (default) vs 192 / 160 threads (5 / 6 solves co-resident per SM)
for v in "" _t192 _t160; do
  lib=toyslam_b200/lib/libndt_b200$v.so
  echo "== $lib"
  NDTB200_LIB=$lib python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sharded 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'results_only',round(d['e2e']['results_only']['value']),'latency_ms',d['latency']['ms_per_align'], 'parity', d.get('parity_vs_oracle'))"
done
