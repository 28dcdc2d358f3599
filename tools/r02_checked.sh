# parity suite on the CHECKED build (-DNDTB200_CHECKED: NDT_CHECK device assertions compiled in); compute-sanitizer cannot
# attach on this GPU pool (gpurun_out/r02_sanitize_*.log)
mkdir -p gpurun_out
NDTB200_LIB=toyslam_b200/lib/libndt_b200_checked.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_full_size.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_checked_pytest.log
NDTB200_LIB=toyslam_b200/lib/libndt_b200_checked.so python tools/sanitize_case.py 2>&1 | tail -2 | tee -a gpurun_out/r02_checked_pytest.log
