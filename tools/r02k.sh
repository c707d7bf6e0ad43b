#!/bin/bash
# round 2, GPU call K: multi-iteration kernel with ordinary coherent loads (parity, small boxes, ncu), then the ncu captures
# behind profiles/traffic.json for the round-2 kernels and the launch list of the default bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -x -q -m gpu -k "several_iterations or golden" > $O/r02k_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02k_pytest.log
tail -3 $O/r02k_pytest.log
B="python bench.py --no-e2e --no-cpu --no-extras --steps 200 --warmup 20"
for w in cavity64 cavity96 cavity128 cavity160; do
  for v in "" "--no-persistent" "--graph-iters 50" "--graph-iters 0"; do
    timeout 300 $B --workload $w $v > $O/r02k_small.json 2> $O/r02k_small.err
    python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02k_small.json").read().strip().splitlines()[-1])
    print("$w [$v]:", round(j["value"]), "MLUPS", round(j["ms_per_step"]*1000,2), "us/step", "frac", round(j["roofline"]["frac"],3), "iters/launch", j["config"]["iterations_per_launch"])
except Exception as e:
    print("$w [$v] FAILED", e, open("gpurun_out/r02k_small.err").read()[-400:])
PY
  done
done 2>&1 | tee $O/r02k_small_sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_multi -s 2 -c 1 -f -o $O/r02k_multi64 \
    python bench.py --workload cavity64 --steps 40 --warmup 20 --no-cpu --no-e2e --no-extras > $O/r02k_ncu_multi64.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_step -s 2 -c 1 -f -o $O/r02k_dense512 \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > $O/r02k_ncu_dense512.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_step -s 2 -c 1 -f -o $O/r02k_slab1024 \
    python bench.py --workload slab1024 --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > $O/r02k_ncu_slab.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_block_step -s 2 -c 1 -f -o $O/r02k_block_sphere \
    python bench.py --workload sphere --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > $O/r02k_ncu_block.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_step -s 2 -c 1 -f -o $O/r02k_dense_q27f64 \
    python bench.py --workload d3q27f64 --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > $O/r02k_ncu_q27.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02k_launches_default.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $O/r02k_ncu_launch_run.log 2>&1
ls -la $O/*.ncu-rep | tail -8
