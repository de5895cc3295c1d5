set -x
for i in 1 2 3; do python -m pytest tests/test_gpu_state.py -q 2>&1 | tail -3; done > gpurun_out/r02c4_state.log 2>&1
for i in 1 2 3; do DKB_FETCH_DMA=1 python -m pytest tests/test_gpu_state.py -q 2>&1 | tail -3; done > gpurun_out/r02c4_state_dma.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02c4_all.log
