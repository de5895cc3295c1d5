set -x
python -m pytest tests/test_gpu_gs.py -x -q > gpurun_out/r02c6_tests.log 2>&1
for m in 4096 1024 2048; do GS_MESH=$m python tools/tuning/time_gs2.py | head -1; done > gpurun_out/r02c6_gs.log 2>&1
GS_MESH=4096 GS_ROWS=512 python tools/tuning/time_gs2.py | head -1 >> gpurun_out/r02c6_gs.log 2>&1
GS_VARIANT=31 GS_MESH=4096 python tools/tuning/time_gs2.py | head -1 >> gpurun_out/r02c6_gs.log 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_state.py tests/test_gpu_multi.py -x -q > gpurun_out/r02c6_tests2.log 2>&1
