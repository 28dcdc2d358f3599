mkdir -p gpurun_out
CMD="python bench.py --steps 128 --warmup 3 --replicas 8 --no-cpu-baseline --e2e-steps 8 --latency-steps 8 --cache /tmp/wl"
timeout 600 $CMD > gpurun_out/p2_plain.json 2> gpurun_out/p2_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01b.csv $CMD > gpurun_out/p2_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ndt_align_kernel -s 12 -c 3 -o gpurun_out/align_r01e -f $CMD > gpurun_out/p2_ncu2.log 2>&1
ls -la gpurun_out
