mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/m4_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/m4_pytest.log
tail -4 gpurun_out/m4_pytest.log
timeout 900 $TR --master-port 29811 bench.py --gpus 4 --no-cpu-baseline > gpurun_out/m4_bench_n4.json 2> gpurun_out/m4.err
timeout 900 $TR --master-port 29812 bench.py --gpus 4 --workload c3 --steps 8192 --no-cpu-baseline > gpurun_out/m4_c3_n4.json 2>> gpurun_out/m4.err
timeout 900 $TR --master-port 29813 bench.py --gpus 4 --workload c4 --c4-city-points 100000000 --steps 30 > gpurun_out/m4_c4_full_n4.json 2>> gpurun_out/m4.err
timeout 900 $TR --master-port 29814 bench.py --gpus 4 --workload c5 --c5-points 500000000 --c5-res 1.0 > gpurun_out/m4_c5_n4.json 2>> gpurun_out/m4.err
tail -3 gpurun_out/m4.err
