set -x
python -m pytest tests/test_gpu_parity.py::test_web_jac_zero_pivot_is_redone_exactly tests/test_gpu_state.py -x -q 2>&1 | tail -8 > gpurun_out/r02c5_a.log
python -m pytest tests/test_gpu_parity.py::test_web_jac_kernel_variants_agree tests/test_gpu_state.py -x -q 2>&1 | tail -8 > gpurun_out/r02c5_b.log
python -m pytest tests/test_gpu_gs.py tests/test_gpu_state.py -x -q 2>&1 | tail -8 > gpurun_out/r02c5_c.log
python -m pytest tests/test_gpu_parity.py tests/test_gpu_state.py -x -q -k "not zero_pivot and not variants_agree" 2>&1 | tail -8 > gpurun_out/r02c5_d.log
for v in 32 33; do for ne in 0 1; do
  if [ $ne = 1 ]; then export DKB_GS_NOEDGE=1; else unset DKB_GS_NOEDGE; fi
  GS_VARIANT=$v GS_MESH=4096 python tools/tuning/time_gs2.py | head -1
  GS_VARIANT=$v GS_MESH=1024 python tools/tuning/time_gs2.py | head -1
done; done > gpurun_out/r02c5_gs.log 2>&1
