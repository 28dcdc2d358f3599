# payload-sort build: parity (bit-identical to the (key, index) path), then the build bench at 10 M / 100 M points
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "payload_sort" 2>&1 | tail -3
python tools/build_bench.py --points 10000000 100000000 --res 1.0 --reps 3 > gpurun_out/r02_payload_build.json 2>&1; cut -c1-200 gpurun_out/r02_payload_build.json
for v in "$@"; do
  echo "== variant $v"
  NDTB200_LIB=toyslam_b200/lib/libndt_b200_pay_$v.so python tools/build_bench.py --points 100000000 --res 1.0 --reps 3 2>&1 | cut -c1-200
done
