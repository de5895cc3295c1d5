set -x
for pc in 512 1024 2048; do DKB_GS_FORCE_PANELS=1 DKB_GS_PANEL_COLS=$pc GS_MESH=4096 python tools/tuning/time_gs2.py | head -1 | sed "s/$/ panel_cols=$pc (+ one 5.4 GB copy back)/"; done > gpurun_out/r02c20_gs.log 2>&1
GS_MESH=4096 python tools/tuning/time_gs2.py | head -1 >> gpurun_out/r02c20_gs.log 2>&1
