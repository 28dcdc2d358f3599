# c3: staged vs fused (cluster launch) builds on throughput-mode handles, 8 / 16 lanes; then the parity tests of the builds
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_and_staged or run_pairs or voxelgrid or mapping" 2>&1 | tail -4
for B in staged default; do
 for L in 8 16; do
  echo "== c3 build path $B lanes $L"
  if [ $B = staged ]; then export NDTB200_BUILD_PATH=staged; else unset NDTB200_BUILD_PATH; fi
  python bench.py --workload c3 --steps 4096 --no-cpu-baseline --c3-lanes $L 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'launches/pair',d['gpu_launches']/d['steps'],'max_err',d['max_translation_error_vs_truth_m'])"
 done
done
unset NDTB200_BUILD_PATH
for C in 8 4; do
  echo "== c3 default path, cluster of $C CTAs, 8 lanes"
  NDTB200_FUSED_CTAS=$C python bench.py --workload c3 --steps 4096 --no-cpu-baseline --c3-lanes 8 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('value',round(d['value']),'launches/pair',d['gpu_launches']/d['steps'])"
done
