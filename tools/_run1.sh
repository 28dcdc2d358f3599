mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fitness" > gpurun_out/t19_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t19_pytest.log
tail -4 gpurun_out/t19_pytest.log
python - <<'PY' > gpurun_out/t19_fitness_timing.log 2>&1
import sys, time, os
sys.path.insert(0, '.')
import numpy as np, bench, toyslam_b200 as nb
class A: pass
a=A(); a.map_points=1000000; a.map_scans=31; a.azimuth_steps=1875; a.cache='/tmp/wl'; a.method='DIRECT7'
w=bench.make_workload(a,0,1)
ndt=nb.NormalDistributionsTransform(); ndt.setInputTarget(w['target']); ndt.setInputSource(w['source']); ndt.align()
os.environ['NDTB200_DEBUG_STEP']='1'
ndt.getFitnessScore()
del os.environ['NDTB200_DEBUG_STEP']
t=time.perf_counter()
for _ in range(5): f=ndt.getFitnessScore()
print('grid', f, (time.perf_counter()-t)/5*1e3, 'ms')
PY
cat gpurun_out/t19_fitness_timing.log
