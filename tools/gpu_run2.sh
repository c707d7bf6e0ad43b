#!/bin/bash
# C++ veneer on the GPU: parity tests, then the benchmark binary itself (same CLI as the reference benchmark).
mkdir -p gpurun_out
O=gpurun_out
APP=neon_b200/cpp/bin/lbm-lid-driven-cavity-flow
timeout 900 python -m pytest tests/test_cpp_veneer.py -x -q -m gpu 2>&1 | tail -25 > $O/pytest_cpp.log
cd $O
B="--grid dGrid --computeFP float --storageFP float --benchmark --warmup-iter 10"
( for n in 64 128 256; do
    timeout 300 ../$APP --deviceType gpu --deviceIds 0 --domain-size $n --max-iter 210 --device-setup $B --report-filename cpp_$n
    timeout 300 ../$APP --deviceType gpu --deviceIds 0 --domain-size $n --max-iter 210 --device-setup --graph $B --report-filename cpp_${n}_graph
  done
  timeout 600 ../$APP --deviceType gpu --deviceIds 0 --domain-size 512 --max-iter 110 $B --report-filename cpp_512
  timeout 600 ../$APP --deviceType gpu --deviceIds 0 0 --domain-size 512 --max-iter 110 --device-setup --sOCC $B --report-filename cpp_512_2parts_socc
  timeout 600 ../$APP --deviceType gpu --deviceIds 0 0 --domain-size 512 --max-iter 110 --device-setup --nOCC --put $B --report-filename cpp_512_2parts_nocc_put
) > cpp_app.log 2>&1
grep -h -A3 "MLUPS\|Problem Setup\|Grid Init" cpp_app.log | grep -v "^--" > cpp_app_metrics.log
