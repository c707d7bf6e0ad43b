#!/bin/bash
# round 2, GPU call B: pipe throughputs the bit-exact collision leans on, full GPU suite (D3Q27 goldens included),
# ncu --set full of the REFERENCE-arithmetic kernel at 512^3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_probe tools/pipe_probe.cu && /tmp/pipe_probe > $O/r02b_pipe_probe.log 2>&1
cat $O/r02b_pipe_probe.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_fast_parity.py > $O/r02b_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02b_pytest.log
tail -8 $O/r02b_pytest.log
B="python bench.py --no-e2e --no-cpu --steps 50 --warmup 5"
timeout 300 $B --arith reference > $O/r02b_bench_ref.json 2> $O/r02b_bench_ref.err
timeout 300 $B --arith reference --workload d3q27f64 > $O/r02b_bench_ref_q27.json 2> $O/r02b_bench_ref_q27.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_step -s 2 -c 1 -f -o $O/r02b_ref512 \
    python bench.py --arith reference --steps 3 --warmup 3 --no-cpu --no-e2e > $O/r02b_ncu_ref512.log 2>&1
python - <<'PY'
import json
for n in ("ref", "ref_q27"):
    try:
        j = json.loads(open(f"gpurun_out/r02b_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["value"]), round(j["ms_per_step"], 4), round(j["roofline"]["frac"], 4))
    except Exception as e:
        print(n, "FAILED", e)
PY
