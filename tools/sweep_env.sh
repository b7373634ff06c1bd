#!/bin/bash
# usage: tools/sweep_env.sh VAR v1 v2 ...   — default bench workload, device phases only
var=$1; shift
for v in "$@"; do
  echo "== $var=$v"
  env $var=$v timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | grep -E "ms/step; phases"
done
