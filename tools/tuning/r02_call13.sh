set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_gs.py -x -q -k "res or datv or fused or paths_agree or single_calls or sampled or solve_web" > gpurun_out/r02c13_tests.log 2>&1
RES_VARIANTS=0,0 python tools/tuning/tune_resid.py > gpurun_out/r02c13_resid.log 2>&1
python tools/tuning/tune_datv.py > gpurun_out/r02c13_datv.log 2>&1
