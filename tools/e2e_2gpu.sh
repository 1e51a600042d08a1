#!/bin/bash
# modeler job end to end on 16 000 config-#3 series (23 M rows of CSV): 1 rank vs 2 ranks, then bench.py on 2 GPUs
set -e
D=/tmp/e2e_scale
rm -rf $D
python tools/e2e_scaling.py make $D 16000 1440
echo "== 1 rank"; timeout 600 python tools/e2e_scaling.py run $D 2>&1 | grep -E "^\{|Error|error" | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
echo "== 2 ranks"; timeout 600 $TR tools/e2e_scaling.py run $D 2>&1 | grep -E "^\{|Error|error" | tail -2
echo "== bench 2 GPUs"; timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | grep -E "^\{|Error|error|Traceback" | tail -3
