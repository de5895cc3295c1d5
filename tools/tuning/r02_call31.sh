set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "jac" > gpurun_out/r02c31_tests.log 2>&1
JAC_VARIANTS=0,10,0 python tools/tuning/tune_jac.py > gpurun_out/r02c31_jac1024.log 2>&1
JAC_MESH=4096 JAC_JPRE=3 JAC_VARIANTS=0,10 python tools/tuning/tune_jac.py > gpurun_out/r02c31_jac4096.log 2>&1
