mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t32_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t32_pytest.log
tail -30 gpurun_out/t32_pytest.log
