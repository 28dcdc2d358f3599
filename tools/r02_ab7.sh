# does aligning c2 from a voxel-key-sorted source help (threshold normally 262144 points)?
for m in 262144 65536; do
  echo "== NDTB200_SORT_SOURCE_MIN=$m"
  NDTB200_SORT_SOURCE_MIN=$m python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sharded 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'results_only',round(d['e2e']['results_only']['value']),'latency_ms',d['latency']['ms_per_align'], 'parity', d.get('parity_vs_oracle'))"
done
