#!/bin/bash
# 2-GPU job: pool threads -> devices from the C++ host, and the torchrun bench at N = 2
nvidia-smi -L > gpurun_out/r2j_gpus.txt 2>&1
python -m pytest tests/test_integration.py -m gpu -q -k "pool_threads" > gpurun_out/r2j_pool.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 examples/global_lighting.py > gpurun_out/r2j_lighting.log 2>&1
tail -3 gpurun_out/r2j_pool.log; tail -c 600 gpurun_out/r2j_bench_n2.json
