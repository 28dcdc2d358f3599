mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t30_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t30_pytest.log
tail -6 gpurun_out/t30_pytest.log
timeout 600 python tools/exp_hess.py > gpurun_out/t30_hess.log 2>&1; cat gpurun_out/t30_hess.log
NDTB200_DEBUG_STEP=1 timeout 600 python tools/timeline.py --cache /tmp/wl > gpurun_out/t30_timeline.log 2>&1
grep -n "step 1:\|eval  1:\|L2: kernel" gpurun_out/t30_timeline.log | head -12
timeout 900 python bench.py --no-cpu-baseline --cache /tmp/wl --e2e-steps 64 > gpurun_out/t30_bench.json 2> gpurun_out/t30_bench.err
