#!/bin/bash
for cfg in "12 4 3" "12 5 3" "13 5 3" "13 4 3" "11 4 3" "12 5 4"; do
  set -- $cfg
  echo "== T=$1 R=$2 low_bits=$3"
  QB_TILE_BITS=$1 QB_REG_BITS=$2 QB_LOW_BITS=$3 python scripts/quick_bench.py 2>&1 | grep -E "circuit"
done
