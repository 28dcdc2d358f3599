mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python bench.py --workload c4 --c4-city-points 100000000 --steps 30 > gpurun_out/m2_c4_full_n1.json 2> gpurun_out/m2.err
timeout 900 $TR --master-port 29801 bench.py --gpus 2 --workload c4 --c4-city-points 100000000 --steps 30 > gpurun_out/m2_c4_full_n2.json 2>> gpurun_out/m2.err
timeout 900 $TR --master-port 29802 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/m2_bench_n2.json 2>> gpurun_out/m2.err
timeout 900 $TR --master-port 29803 bench.py --gpus 2 --workload c3 --steps 4096 > gpurun_out/m2_c3_n2.json 2>> gpurun_out/m2.err
timeout 900 $TR --master-port 29804 bench.py --gpus 2 --workload c5 --c5-points 100000000 500000000 --c5-res 1.0 > gpurun_out/m2_c5_n2.json 2>> gpurun_out/m2.err
timeout 900 $TR --master-port 29805 bench.py --gpus 2 --workload c4 --steps 50 --map-points 4000000 --map-scans 61 --thin-leaf 0.07 > gpurun_out/m2_c4_scaled_n2.json 2>> gpurun_out/m2.err
tail -5 gpurun_out/m2.err
