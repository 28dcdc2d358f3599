# last call of the round: full GPU suite + the driver's two bench invocations + build sweep
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_final_pytest.log; cat gpurun_out/r02_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; tail -1 gpurun_out/r02_final_smoke.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_final_reference_arm.json 2> gpurun_out/r02_final.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_final_bench_n1.json 2>> gpurun_out/r02_final.err
python bench.py --workload c5 --c5-points 10000000 100000000 500000000 > gpurun_out/r02_final_bench_c5.json 2>> gpurun_out/r02_final.err
python bench.py --workload c1 > gpurun_out/r02_final_bench_c1.json 2>> gpurun_out/r02_final.err
tail -c 300 gpurun_out/r02_final.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_final_bench_n1.json') if l.startswith('{')][-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'lat',d['latency']['ms_per_align'],'ref',json.loads([l for l in open('gpurun_out/r02_final_reference_arm.json') if l.startswith('{')][-1])['value'])
for l in open('gpurun_out/r02_final_bench_c5.json'):
    if l.startswith('{'):
        d=json.loads(l); print('c5',d['config']['workload'][30:75], round(d['ms_per_step'],3), round(d['roofline']['frac'],4))
d=json.loads([l for l in open('gpurun_out/r02_final_bench_c1.json') if l.startswith('{')][-1])
print('c1',{k:(round(v['single_ms'],3),round(v['ktimes_ms'],3),v['fitness_matches_readme']) for k,v in d['methods'].items()})
PY
