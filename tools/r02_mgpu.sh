# Round-2 multi-GPU records (N = 2 / 4 / 8 GPUs of one box): bash tools/r02_mgpu.sh N [full]
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r02_m${N}_pytest.log; cat gpurun_out/r02_m${N}_pytest.log
$TR --master-port 29811 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_m${N}_bench_n${N}.json 2> gpurun_out/r02_m${N}.err; tail -c 300 gpurun_out/r02_m${N}.err
$TR --master-port 29813 bench.py --gpus $N --workload c3 --steps 8192 --no-cpu-baseline > gpurun_out/r02_m${N}_c3_n${N}.json 2>> gpurun_out/r02_m${N}.err
if [ "$2" = "full" ]; then  # the cluster-launched fused build next to the default (staged) one
  NDTB200_BUILD_PATH=fused $TR --master-port 29814 bench.py --gpus $N --workload c3 --steps 8192 --no-cpu-baseline > gpurun_out/r02_m${N}_c3_cluster_n${N}.json 2>> gpurun_out/r02_m${N}.err
fi
if [ "$2" = "full" ]; then
  $TR --master-port 29812 bench.py --gpus $N --workload c5 --c5-points 100000000 --c5-res 1.0 > gpurun_out/r02_m${N}_c5_n${N}.json 2>> gpurun_out/r02_m${N}.err
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_m${N}_*.json")):
    for line in open(f):
        if line.startswith("{"):
            d=json.loads(line)
            print(f, "value", round(d["value"]), d.get("unit"), "e2e", d.get("e2e",{}).get("value"), "sharded", (d.get("sharded") or {}).get("ms_per_align"), (d.get("sharded") or {}).get("roofline",{}).get("frac"))
PY
