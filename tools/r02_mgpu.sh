N=${1:-2}
python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err; tail -c 600 gpurun_out/r02_bench_n${N}.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus $N --workload c5 --c5-points 100000000 --c5-res 1.0 > gpurun_out/r02_c5_n${N}.json 2> gpurun_out/r02_c5_n${N}.err; tail -c 400 gpurun_out/r02_c5_n${N}.err
python bench.py --workload c5 --c5-points 100000000 --c5-res 1.0 > gpurun_out/r02_c5_n1_ref.json 2>/dev/null
echo done
