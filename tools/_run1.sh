mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_bench.py -m gpu -x -q > gpurun_out/t31_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t31_pytest.log
tail -25 gpurun_out/t31_pytest.log
