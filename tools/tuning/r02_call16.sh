set -x
for v in 0 41; do for ns in 20 100 400; do
  DKB_GS_POLL_NS=$ns GS_VARIANT=$v GS_MESH=4096 python tools/tuning/time_gs2.py | head -1 | sed "s/$/ poll_ns=$ns/"
done; done > gpurun_out/r02c16_gs.log 2>&1
