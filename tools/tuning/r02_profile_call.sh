set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02s2_tests.log 2>&1
python bench.py > gpurun_out/r02s2_bench_n1.json 2> gpurun_out/r02s2_bench_n1.err
python bench.py --steps 2 --warmup 2 --no-cpu --jpre2 0 > gpurun_out/r02s2_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 2 --no-cpu --jpre2 0 > gpurun_out/r02s2_ncu1.log 2>&1
python bench.py --mx 2048 --my 2048 --steps 2 --warmup 2 --no-cpu --jpre2 0 > gpurun_out/r02s2_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'tpp_psol|web_datv_single|reduce_kernel|web_jac2' -s 150 -c 10 -o gpurun_out/r02_top python bench.py --mx 2048 --my 2048 --steps 2 --warmup 2 --no-cpu --jpre2 0 > gpurun_out/r02s2_ncu2.log 2>&1
GS_MESH=4096 python tools/tuning/time_gs2.py > gpurun_out/r02s2_gs_plain.log 2>&1 && GS_MESH=4096 ncu --set full --clock-control none --import-source on -k regex:gs_stream -s 3 -c 1 -o gpurun_out/r02_gs_stream python tools/tuning/time_gs2.py > gpurun_out/r02s2_ncu3.log 2>&1
