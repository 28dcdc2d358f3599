mkdir -p gpurun_out
CMD="python bench.py --steps 256 --warmup 3 --replicas 8 --no-cpu-baseline --e2e-steps 8 --latency-steps 16 --cache /tmp/wl"
timeout 600 $CMD > gpurun_out/p3_plain.json 2> gpurun_out/p3_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01c.csv $CMD > gpurun_out/p3_ncu1.log 2>&1
# full capture: two launches of the latency shape (1024 threads) and two of the throughput shape (256 threads)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ndt_align_kernel -s 12 -c 2 -o gpurun_out/align_r01_lat -f $CMD > gpurun_out/p3_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ndt_align_kernel -s 60 -c 2 -o gpurun_out/align_r01_tp -f $CMD > gpurun_out/p3_ncu3.log 2>&1
# map build: launch list + full capture of the scatter and moments kernels on a 10 M-point cloud
BCMD="python tools/build_bench.py --points 10000000 --res 1.0 --reps 1"
timeout 600 $BCMD > gpurun_out/p3_build_plain.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r01_build.csv $BCMD > gpurun_out/p3_ncu4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"radix_scatter|voxel_moments" -s 3 -c 4 -o gpurun_out/build_r01 -f $BCMD > gpurun_out/p3_ncu5.log 2>&1
ls -la gpurun_out | tail -20
