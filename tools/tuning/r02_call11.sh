set -x
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_gs.py -x -q -k "bit_exact and stream and 10-100-70" > gpurun_out/r02c11_racecheck.log 2>&1
JAC_VARIANTS=0,4,0,4 python tools/tuning/tune_jac.py > gpurun_out/r02c11_jac.log 2>&1
python bench.py --workload lu > gpurun_out/r02c11_lu.json 2> gpurun_out/r02c11_lu.err
python bench.py --workload heat > gpurun_out/r02c11_heat.json 2> gpurun_out/r02c11_heat.err
