set -x
python -m pytest tests/test_gpu_gs.py tests/test_gpu_parity.py tests/test_gpu_state.py tests/test_gpu_roots.py -x -q > gpurun_out/r02c3_tests.log 2>&1
GS_MESH=4096 python tools/tuning/time_gs2.py > gpurun_out/r02c3_gs.log 2>&1
GS_MESH=1024 python tools/tuning/time_gs2.py >> gpurun_out/r02c3_gs.log 2>&1
GS_MESH=4096 GS_ROWS=512 python tools/tuning/time_gs2.py >> gpurun_out/r02c3_gs.log 2>&1
python tools/tuning/tune_resid.py > gpurun_out/r02c3_resid.log 2>&1
python bench.py --no-cpu > gpurun_out/r02c3_bench.json 2> gpurun_out/r02c3_bench.err
