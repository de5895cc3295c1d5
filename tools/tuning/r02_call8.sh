set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02c8_tests.log 2>&1
python bench.py > gpurun_out/r02c8_bench.json 2> gpurun_out/r02c8_bench.err
python bench.py --workload web1024 --no-cpu > gpurun_out/r02c8_bench1024.json 2> gpurun_out/r02c8_bench1024.err
( time python bench.py --impl reference ) > gpurun_out/r02c8_ref.json 2> gpurun_out/r02c8_ref.err
