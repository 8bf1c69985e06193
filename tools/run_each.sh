#!/bin/bash
# run every test of a file in its own process (a CUDA fault poisons the context)
f=$1; out=$2; : > $out
for t in $(python -m pytest $f --collect-only -q 2>/dev/null | grep "::"); do
  echo "=== $t" >> $out
  timeout 300 python -m pytest "$t" -x -q 2>&1 | tail -25 >> $out
done
