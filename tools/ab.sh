#!/bin/bash
# usage: tools/ab.sh "<variants>"  -- c3 50k timing for each build variant
for v in $1; do
  vv=$v; [ "$v" = "base" ] && vv=""
  echo "== variant $v"; PB200_VARIANT=$vv timeout 200 python tools/prof_fit.py 50000 2 c3 2>&1 | tail -1
  PB200_VARIANT=$vv timeout 100 python tools/prof_fit.py 200000 2 c4 2>&1 | tail -1
done
