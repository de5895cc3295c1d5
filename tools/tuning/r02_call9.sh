set -x
python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r02c9_tests.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --no-cpu > gpurun_out/r02c9_bench_n2.json 2> gpurun_out/r02c9_bench_n2.err
