#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for m in step chain chain_all chain_nospec chain_nokeep; do  # (coop: the resident-grid kernel, retired after this run)
  echo "=== $m"
  timeout 300 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 2 python tools/race_probe.py $m 2>&1 | grep -v "^=========$" | cut -c1-300 | tail -14
done > $O/r02q_race.log 2>&1
cat $O/r02q_race.log
