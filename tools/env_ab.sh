#!/bin/bash
# A/B of environment settings in ONE call on ONE box: cold benchmark generation, constant 72 towers (tools/phase_ab.py
# --only-tables), alternating processes.  usage: env_ab.sh "VAR=a" "VAR=b OTHER=c" ...   ("" = defaults)
mkdir -p gpurun_out
for i in 1 2; do
  for e in "" "$@"; do
    echo "env='${e}' $(env $e timeout 100 python tools/phase_ab.py --reps 1 --only-tables 2>&1 | grep '^tables 1' | awk '{printf "%s %s %s | ", $3, $5, $6}')"
  done
done
