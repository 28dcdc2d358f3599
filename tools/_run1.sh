mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t28_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t28_pytest.log
tail -12 gpurun_out/t28_pytest.log
timeout 900 python bench.py --workload mapper --steps 200 > gpurun_out/t28_mapper.json 2> gpurun_out/t28_mapper.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t28_smoke.log 2>&1; tail -2 gpurun_out/t28_smoke.log
