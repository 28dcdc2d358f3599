mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/f_pytest.log
tail -5 gpurun_out/f_pytest.log
timeout 600 python bench.py --workload mapper --steps 200 > gpurun_out/r01c_bench_mapper.json 2> gpurun_out/f.err
timeout 900 python bench.py --impl reference > gpurun_out/r01c_bench_reference_arm.json 2>> gpurun_out/f.err
timeout 900 python bench.py > gpurun_out/r01c_bench_n1.json 2>> gpurun_out/f.err
timeout 900 python bench.py --method DIRECT1 --no-cpu-baseline --e2e-steps 64 --latency-steps 50 > gpurun_out/r01c_bench_direct1.json 2>> gpurun_out/f.err
timeout 900 python bench.py --method DIRECT26 --no-cpu-baseline --e2e-steps 64 --latency-steps 50 > gpurun_out/r01c_bench_direct26.json 2>> gpurun_out/f.err
timeout 900 python bench.py --workload c3 --steps 4096 > gpurun_out/r01c_bench_c3.json 2>> gpurun_out/f.err
timeout 900 python bench.py --workload c5 --c5-points 10000000 100000000 500000000 > gpurun_out/r01c_bench_c5.json 2>> gpurun_out/f.err
timeout 600 python tools/timeline.py > gpurun_out/r01c_timeline.txt 2>&1
tail -3 gpurun_out/f.err
