set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 8 --no-cpu > gpurun_out/r02c21_bench_n8.json 2> gpurun_out/r02c21_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 4 --no-cpu > gpurun_out/r02c21_bench_n4.json 2> gpurun_out/r02c21_bench_n4.err
