set -x
C4="python bench.py --workload c4 --c4-city-points 100000000 --steps 3 --warmup 3"
$C4 > gpurun_out/r02_c4_n1_before.json 2> gpurun_out/r02_c4_n1_before.err && \
timeout 700 ncu --set full --clock-control none --import-source on -k regex:ndt_align_kernel -s 2 -c 1 -f -o gpurun_out/r02_c4_before $C4 > gpurun_out/r02_c4_ncu.log 2>&1
BB="python tools/build_bench.py --points 100000000 --res 1.0 --reps 1"
$BB > gpurun_out/r02_build100m_before.json 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ndtb200 --csv --log-file gpurun_out/r02_build100m_launches_before.csv $BB > gpurun_out/r02_build100m_ncu.log 2>&1
cat gpurun_out/r02_c4_n1_before.json gpurun_out/r02_build100m_before.json
