#!/bin/bash
echo "== c4 n=4096"; timeout 60 python tools/prof_fit.py 4096 2 c4 2>&1 | tail -2
echo "== c4 n=65536"; timeout 90 python tools/prof_fit.py 65536 2 c4 2>&1 | tail -2
echo "== memcheck c4 n=256"; timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/prof_fit.py 256 1 c4 2>&1 | tail -15
echo "== memcheck c3 n=64"; timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/prof_fit.py 64 1 c3 2>&1 | tail -8
echo "== c2"; timeout 60 python tools/prof_fit.py 1000 2 c2 2>&1 | tail -2
