#!/bin/bash
# round 2, GPU call U (2 GPUs): the multi-GPU tests and the default bench line at N=2 with the final library
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multiproc.py tests/test_cpp_veneer.py -q -m gpu > $O/r02u_pytest_multi2.log 2>&1
echo "pytest rc=$?" >> $O/r02u_pytest_multi2.log
tail -3 $O/r02u_pytest_multi2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02u_bench2.json 2> $O/r02u_bench2.err; echo "bench rc=$?"
timeout 600 $TR --master-port 29522 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $O/r02u_bench2_reference.json 2> $O/r02u_bench2_reference.err; echo "ref rc=$?"
python - <<PY
import json
j=json.loads(open("gpurun_out/r02u_bench2.json").read().strip().splitlines()[-1])
print("headline", round(j["value"]), j["ms_per_step"], j["roofline"]["frac"], "e2e", j["e2e"] and round(j["e2e"]["value"]), "arith_ref", j["arith_reference"] and round(j["arith_reference"]["value"]))
for e in j["extra_configs"]:
    print(" ", e.get("key"), e.get("error") or (round(e["value"]), round(e["roofline"]["frac"],3), e.get("issue")))
PY
tail -c 300 $O/r02u_bench2_reference.json
