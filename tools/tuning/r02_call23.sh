set -x
for v in 0 42 0 42; do GS_VARIANT=$v GS_MESH=4096 python tools/tuning/time_gs2.py | head -1; done > gpurun_out/r02c23_gs.log 2>&1
