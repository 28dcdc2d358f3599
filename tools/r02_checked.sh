echo "== parity suite on the CHECKED build (NDT_CHECK assertions compiled in)"
NDTB200_LIB=toyslam_b200/lib/libndt_b200_checked.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_full_size.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_checked_pytest.log
NDTB200_LIB=toyslam_b200/lib/libndt_b200_checked.so python tools/sanitize_case.py 2>&1 | tail -2 | tee -a gpurun_out/r02_checked_pytest.log
echo "== c3 / mapper after the host-side changes"
python bench.py --workload c3 --steps 4096 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('c3 value',d['value'],'launches/pair',d['gpu_launches']/d['steps'])"
python bench.py --workload mapper --steps 200 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('mapper value',d['value'],'ms_per_step',d['ms_per_step'], d['step_ms_percentiles'])"
