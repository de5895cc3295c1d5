set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "jac" > gpurun_out/r02c2_tests.log 2>&1
python tools/tuning/tune_jac.py > gpurun_out/r02c2_jac.log 2>&1 && JAC_VARIANTS=0 ncu --set full --clock-control none --import-source on -k regex:web_jac3 -s 2 -c 1 -o gpurun_out/r02_jac3 python tools/tuning/tune_jac.py > gpurun_out/r02c2_ncu.log 2>&1
JAC_MESH=4096 JAC_JPRE=3 JAC_VARIANTS=0,10 python tools/tuning/tune_jac.py > gpurun_out/r02c2_jac4096.log 2>&1
python tools/tuning/pcie_probe.py > gpurun_out/r02c2_pcie.log 2>&1
