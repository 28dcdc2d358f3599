mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t34_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t34_pytest.log
tail -5 gpurun_out/t34_pytest.log
timeout 600 python tools/build_bench.py --points 30000 60000 --res 1.0 --reps 5 > gpurun_out/t34_build_small.log 2>&1
NDTB200_BUILD_PATH=staged timeout 600 python tools/build_bench.py --points 30000 60000 --res 1.0 --reps 5 > gpurun_out/t34_build_small_staged.log 2>&1
timeout 900 python bench.py --workload c3 --steps 4096 --cache /tmp/wl --no-cpu-baseline > gpurun_out/t34_c3.json 2> gpurun_out/t34.err
timeout 900 python bench.py --workload mapper --steps 200 --cache /tmp/wl --no-cpu-baseline > gpurun_out/t34_mapper.json 2>> gpurun_out/t34.err
NDTB200_BUILD_PATH=staged timeout 900 python bench.py --workload mapper --steps 200 --cache /tmp/wl --no-cpu-baseline > gpurun_out/t34_mapper_staged.json 2>> gpurun_out/t34.err
tail -3 gpurun_out/t34.err
