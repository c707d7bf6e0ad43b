#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multiproc.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest.log
CUDA_LAUNCH_BLOCKING=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_debug.py ipc > gpurun_out/dbg_ipc.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tests/mgpu_check.py > gpurun_out/mgpu2.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --transport ipc > gpurun_out/bench2_ipc.json 2> gpurun_out/bench2_ipc.err
