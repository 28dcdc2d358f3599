mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "8" > gpurun_out/m8_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/m8_pytest.log
tail -4 gpurun_out/m8_pytest.log
timeout 600 $TR --master-port 29821 bench.py --gpus 8 --no-cpu-baseline --steps 2048 --e2e-steps 64 --latency-steps 20 > gpurun_out/m8_bench_n8.json 2> gpurun_out/m8.err
timeout 600 $TR --master-port 29823 bench.py --gpus 8 --workload c4 --c4-city-points 100000000 --steps 30 > gpurun_out/m8_c4_full_n8.json 2>> gpurun_out/m8.err
tail -3 gpurun_out/m8.err | cut -c1-300
