N=${1:-2}
NDTB200_SHARD_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus $N --workload c5 --c5-points 100000000 --c5-res 1.0 2>&1 | grep -E "SHARD_TIMING|^{" | tail -6 | cut -c1-700
