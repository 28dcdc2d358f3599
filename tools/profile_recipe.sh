mkdir -p gpurun_out
CMD="python bench.py --steps 128 --warmup 3 --replicas 8 --no-cpu-baseline --e2e-steps 8 --latency-steps 8 --cache /tmp/wl"
timeout 600 $CMD > gpurun_out/p4_plain.json 2> gpurun_out/p4_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r01c.csv $CMD > gpurun_out/p4_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ndt_align_kernel -s 10 -c 2 -o gpurun_out/align_r01c_lat -f $CMD > gpurun_out/p4_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ndt_align_kernel -s 40 -c 2 -o gpurun_out/align_r01c_tp -f $CMD > gpurun_out/p4_ncu3.log 2>&1
BCMD="python tools/build_bench.py --points 30000 --res 1.0 --reps 1"
timeout 600 $BCMD > gpurun_out/p4_build_plain.json 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:small_build_kernel -c 2 -o gpurun_out/small_build_r01c -f $BCMD > gpurun_out/p4_ncu4.log 2>&1
ls -la gpurun_out | grep r01c
