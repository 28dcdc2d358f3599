for L in "" toyslam_b200/lib/libndt_b200_noprefetch.so; do
  echo "== lib=$L"
  NDTB200_LIB=$L python bench.py --workload c4 --c4-city-points 100000000 --steps 20 --warmup 3 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('c4 ms',d['ms_per_step'],'frac',d['roofline']['frac'])"
done
