#!/bin/bash
# round 2, GPU call R: after retiring the cooperative kernel — sanitizer, chain tests (Python + C++ app), whole suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > $O/r02r_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02r_pytest.log
tail -4 $O/r02r_pytest.log
for n in 96 128; do
 ./neon_b200/cpp/bin/lbm-lid-driven-cavity-flow --deviceType gpu --deviceIds 0 --grid dGrid --domain-size $n --warmup-iter 100 --max-iter 1100 --benchmark --report-filename $O/r02r_app$n 2>&1 | grep -i "mlups"
 ./neon_b200/cpp/bin/lbm-lid-driven-cavity-flow --deviceType gpu --deviceIds 0 --grid dGrid --domain-size $n --warmup-iter 100 --max-iter 1100 --benchmark --graph --report-filename $O/r02r_app${n}g 2>&1 | grep -i "mlups"
done 2>&1 | tee $O/r02r_cpp_app_small.log
