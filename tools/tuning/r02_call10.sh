set -x
python -m pytest tests/test_gpu_multi.py -x -q -k "8" > gpurun_out/r02c10_tests.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --no-cpu > gpurun_out/r02c10_bench_n8.json 2> gpurun_out/r02c10_bench_n8.err
