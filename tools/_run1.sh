mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fitness or mapping" > gpurun_out/t27_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t27_pytest.log
tail -3 gpurun_out/t27_pytest.log
timeout 900 python bench.py --workload mapper --steps 200 --no-cpu-baseline > gpurun_out/t27_mapper.json 2> gpurun_out/t27_mapper.err
