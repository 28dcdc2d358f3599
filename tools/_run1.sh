mkdir -p gpurun_out
L=gpurun_out/t12.log
for v in default sort16x3 sort8x4 sort8x2; do
  if [ $v != default ]; then export NDTB200_LIB=$PWD/build/variants/libndt_$v.so; fi
  echo "== $v" >> $L
  timeout 600 python tools/build_bench.py --points 100000000 --res 1.0 --reps 3 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    try:
        d=json.loads(l); print(d['points'], d['resolution'], 'ms %.3f'%d['ms'])
    except Exception: print(l.strip()[:200])" >> $L
done
unset NDTB200_LIB
timeout 900 python bench.py --workload c3 --steps 1024 --cache /tmp/wl > gpurun_out/t12_c3.json 2> gpurun_out/t12_c3.err
tail -3 gpurun_out/t12_c3.err
timeout 900 python bench.py --workload c3 --steps 1024 --cache /tmp/wl --c3-lanes 16 --no-cpu-baseline > gpurun_out/t12_c3_l16.json 2>> gpurun_out/t12_c3.err
timeout 900 python bench.py --workload c3 --steps 1024 --cache /tmp/wl --c3-lanes 1 --no-cpu-baseline > gpurun_out/t12_c3_l1.json 2>> gpurun_out/t12_c3.err
