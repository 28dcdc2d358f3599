# build parity (all build paths bit-identical, oracle at 10 M points), then the build bench, variants of voxel_build_kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -3
python tools/build_bench.py --points 1000000 10000000 100000000 --res 1.0 --reps 3 > gpurun_out/r02_payload_build.json 2>&1; cut -c1-200 gpurun_out/r02_payload_build.json
for v in "$@"; do
  echo "== variant $v"
  NDTB200_LIB=toyslam_b200/lib/libndt_b200_pay_$v.so python tools/build_bench.py --points 1000000 100000000 --res 1.0 --reps 3 2>&1 | cut -c1-200
done
