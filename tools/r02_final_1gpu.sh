# Round-2 end-of-round measurements on ONE B200 (plain runs; the ncu captures are tools/profile_recipe_r02.sh).
# usage: bash tools/r02_final_1gpu.sh [notests]
mkdir -p gpurun_out
if [ "$1" != "notests" ]; then
  python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_final_pytest.log; cat gpurun_out/r02_final_pytest.log
fi
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_final_reference_arm.json 2> gpurun_out/r02_final.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_final_bench_n1.json 2>> gpurun_out/r02_final.err
python bench.py > gpurun_out/r02_final_bench_n1_default.json 2>> gpurun_out/r02_final.err
python bench.py --method DIRECT1 --no-sharded --no-cpu-baseline > gpurun_out/r02_final_bench_direct1.json 2>> gpurun_out/r02_final.err
python bench.py --method DIRECT26 --no-sharded --no-cpu-baseline > gpurun_out/r02_final_bench_direct26.json 2>> gpurun_out/r02_final.err
python bench.py --workload c1 > gpurun_out/r02_final_bench_c1.json 2>> gpurun_out/r02_final.err
python bench.py --workload c3 > gpurun_out/r02_final_bench_c3.json 2>> gpurun_out/r02_final.err
python bench.py --workload mapper > gpurun_out/r02_final_bench_mapper.json 2>> gpurun_out/r02_final.err
python bench.py --workload c4 --c4-city-points 100000000 --steps 3 --warmup 3 > gpurun_out/r02_final_bench_c4.json 2>> gpurun_out/r02_final.err
python bench.py --workload c5 --c5-points 10000000 100000000 500000000 > gpurun_out/r02_final_bench_c5.json 2>> gpurun_out/r02_final.err
python tools/timeline.py > gpurun_out/r02_final_timeline.txt 2>> gpurun_out/r02_final.err
python bench.py --workload c3 --c3-lanes 4 --no-cpu-baseline > gpurun_out/r02_final_bench_c3_lanes4.json 2>> gpurun_out/r02_final.err
tail -c 600 gpurun_out/r02_final.err
ls -la gpurun_out
