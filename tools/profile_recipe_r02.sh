# Round-2 profile recipe (1 GPU).  Every ncu run follows a plain run of the same command that exited 0.  gpurun pulls back
# at most 64 MiB per call, so the captures are split into parts:  bash tools/profile_recipe_r02.sh align|c4|build
mkdir -p gpurun_out
part=${1:-align}
CMD="python bench.py --steps 2 --warmup 3 --replicas 8 --no-cpu-baseline --no-sharded --e2e-steps 1 --latency-steps 8 --cache /tmp/wl"
C4="python bench.py --workload c4 --c4-city-points 100000000 --steps 3 --warmup 3"
BB="python tools/build_bench.py --points 100000000 --res 1.0 --reps 1"
if [ "$part" = "align" ]; then
  timeout 600 $CMD > gpurun_out/r02_prof_plain.json 2> gpurun_out/r02_prof_plain.err || exit 1
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu1.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:ndt_align_kernel -s 10 -c 1 -o gpurun_out/r02_align_lat -f $CMD > gpurun_out/r02_ncu2.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:ndt_align_kernel -s 32 -c 1 -o gpurun_out/r02_align_tp -f $CMD > gpurun_out/r02_ncu3.log 2>&1
elif [ "$part" = "c4" ]; then
  timeout 600 $C4 > gpurun_out/r02_c4_plain.json 2> gpurun_out/r02_c4_plain.err && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:ndt_align_kernel -s 2 -c 1 -o gpurun_out/r02_c4_final -f $C4 > gpurun_out/r02_ncu4.log 2>&1
elif [ "$part" = "build" ]; then
  timeout 600 $BB > gpurun_out/r02_build_plain.json 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_build_launches.csv $BB > gpurun_out/r02_ncu5.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${BUILD_KERNELS:-onesweep_kernel|voxel_build_kernel}" -s ${BUILD_SKIP:-3} -c ${BUILD_COUNT:-2} -o gpurun_out/r02_build_kernels -f $BB > gpurun_out/r02_ncu6.log 2>&1
fi
du -sh gpurun_out; ls -la gpurun_out | grep r02_
