BB="python tools/build_bench.py --points 100000000 --res 1.0 --reps 1"
$BB > gpurun_out/r02_build100m_plain.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_build100m_launches.csv $BB > gpurun_out/r02_build100m_ncu.log 2>&1
tail -2 gpurun_out/r02_build100m_plain.json
