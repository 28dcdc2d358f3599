mkdir -p gpurun_out
for v in default nopipe; do
  if [ $v != default ]; then export NDTB200_LIB=$PWD/build/variants/libndt_$v.so; fi
  timeout 600 python bench.py --no-cpu-baseline --cache /tmp/wl --e2e-steps 64 > gpurun_out/t14_bench_$v.json 2> gpurun_out/t14_bench_$v.err
done
