set -x
python -m pytest tests -m gpu -x -q -k "sorted_by_voxel or emulated or fixture_a" 2>&1 | tail -5
C4="python bench.py --workload c4 --c4-city-points 100000000 --steps 3 --warmup 3"
$C4 > gpurun_out/r02_c4_n1_sorted.json 2> gpurun_out/r02_c4_n1_sorted.err && \
timeout 700 ncu --set full --clock-control none --import-source on -k regex:ndt_align_kernel -s 2 -c 1 -f -o gpurun_out/r02_c4_sorted $C4 > gpurun_out/r02_c4_ncu2.log 2>&1
cat gpurun_out/r02_c4_n1_sorted.json
