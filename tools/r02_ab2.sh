python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err; tail -c 800 gpurun_out/r02_bench_n1_a.err
python bench.py --workload c1 > gpurun_out/r02_bench_c1_a.json 2>> gpurun_out/r02_bench_n1_a.err
echo "== nocached"
NDTB200_LIB=toyslam_b200/lib/libndt_b200_nocached.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_nocached.json 2>/dev/null
echo done
