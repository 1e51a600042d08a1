#!/bin/bash
for v in $1; do
  vv=$v; [ "$v" = "base" ] && vv=""
  echo "== variant $v"; PB200_VARIANT=$vv timeout 100 python tools/prof_fit.py 200000 3 c4 2>&1 | tail -1
done
