# payload-sort build: parity (bit-identical to the (key, index) path), then the build bench at 10 M / 100 M points, both paths
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "payload_sort or fused_and_staged or map_build" 2>&1 | tail -15 > gpurun_out/r02_payload_pytest.log; cat gpurun_out/r02_payload_pytest.log
python -m pytest tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -5
python tools/build_bench.py --points 10000000 100000000 --res 1.0 --reps 3 > gpurun_out/r02_payload_build.json 2>&1; cat gpurun_out/r02_payload_build.json
NDTB200_PAYLOAD_SORT_MIN=99999999999 python tools/build_bench.py --points 10000000 100000000 --res 1.0 --reps 3 > gpurun_out/r02_pairs_build.json 2>&1; cat gpurun_out/r02_pairs_build.json
