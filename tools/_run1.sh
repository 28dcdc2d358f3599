mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t4_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t4_pytest.log
tail -15 gpurun_out/t4_pytest.log
NDTB200_DEBUG_STEP=1 timeout 600 python tools/timeline.py --cache /tmp/wl > gpurun_out/t4_timeline.log 2>&1
timeout 600 python bench.py --steps 500 --warmup 20 --no-cpu-baseline --e2e-steps 50 --cache /tmp/wl > gpurun_out/t4_bench.json 2> gpurun_out/t4_bench.err
