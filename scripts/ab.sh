#!/bin/bash
# A/B of planner / kernel knobs on the benchmark circuit: scripts/ab.sh "QB_ROT=0" "QB_ROT=1 QB_LITE=0" ...
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python scripts/quick_bench.py 2>&1 | grep -E "circuit|gate pass"
done
