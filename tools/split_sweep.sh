#!/bin/bash
# split engine: cold generation at several pool sizes and tower-CTA counts (compare profiles/README.md crossover table)
for g in 256 1024 2048 8192; do
  timeout 120 python tools/split_check.py --perf-only --games $g --ctas 40,56,72 2>&1 | grep "^split"
done
