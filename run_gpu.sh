#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multiproc.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_mp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tests/mgpu_check.py > gpurun_out/mgpu2.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --transport fused > gpurun_out/bench2_fused.json 2> gpurun_out/bench2_fused.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
