#!/bin/bash
# usage: tools/sweep.sh  -- timing of the c3 fit under different CTA widths
for cfg in "1073741824 1073741824" "0 1073741824" "0 0"; do
  set -- $cfg
  echo "== LC0_MAX=$1 LC1_MAX=$2"
  PB200_LC0_MAX=$1 PB200_LC1_MAX=$2 python tools/prof_fit.py 50000 2 c3 2>&1 | tail -1
done
PB200_LC0_MAX=1073741824 python tools/prof_fit.py 200000 2 c4 2>&1 | tail -1
PB200_LC0_MAX=1073741824 python tools/prof_fit.py 1000 2 c2 2>&1 | tail -1
