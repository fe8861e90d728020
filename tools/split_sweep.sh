#!/bin/bash
# split engine: cold generation at several pool sizes and tower-CTA counts (compare profiles/README.md crossover table)
for g in "$@"; do
  timeout 200 python tools/split_check.py --perf-only --games $g --ctas ${CTAS:-72,40,56,72,88} 2>&1 | grep "^split" | tail -n +2
done
