mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t6_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t6_pytest.log
tail -15 gpurun_out/t6_pytest.log
timeout 900 python bench.py --no-cpu-baseline --cache /tmp/wl > gpurun_out/t6_bench.json 2> gpurun_out/t6_bench.err
tail -3 gpurun_out/t6_bench.err
