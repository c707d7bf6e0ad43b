#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 neon_b200/cpp/bin/generic-containers --deviceIds 0 0 0 --n 36 > $O/generic.log 2>&1
timeout 300 neon_b200/cpp/bin/generic-containers --deviceIds 0 --bench 256 >> $O/generic.log 2>&1
timeout 300 neon_b200/cpp/bin/generic-containers --deviceIds 0 --bench 512 >> $O/generic.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > $O/pytest_gpu_all.log
timeout 600 python bench.py > $O/bench1b.json 2> $O/bench1b.err
