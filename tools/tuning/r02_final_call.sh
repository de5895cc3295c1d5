# final single-GPU validation of the round: the driver's own sequence + the evidence run
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1
python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
python bench.py --workload web1024 --no-cpu > gpurun_out/r02f_bench1024.json 2> gpurun_out/r02f_bench1024.err
bash tools/tuning/r02_profile_call.sh
