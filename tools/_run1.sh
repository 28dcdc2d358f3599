mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t21_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t21_pytest.log
tail -30 gpurun_out/t21_pytest.log
