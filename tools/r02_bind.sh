# rank -> NUMA-node core binding check at N GPUs: default vs NDTB200_NO_BIND=1
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for B in bind nobind; do
  if [ $B = nobind ]; then export NDTB200_NO_BIND=1; else unset NDTB200_NO_BIND; fi
  $TR --master-port 29821 bench.py --gpus $N --steps 20 --warmup 3 --no-sharded 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$B value',round(d['value']),'e2e',round(d['e2e']['value']),'results_only',round(d['e2e']['results_only']['value']),'topology',d.get('topology'))"
done
unset NDTB200_NO_BIND
$TR --master-port 29822 bench.py --gpus $N --workload c3 --steps 8192 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('c3 value',round(d['value']),d['config']['workload'][-60:],'topology',d.get('topology'))"
nvidia-smi topo -m 2>/dev/null | head -14
