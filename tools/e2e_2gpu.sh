#!/bin/bash
set -e
D=/tmp/e2e_run
rm -rf $D
python tools/e2e_jobs.py make $D 2000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR -m time_series_spark_b200.modeler_driver $D/modeler.yaml 2>&1 | grep -E "Modeling|Output df|Error|error" | head
timeout 300 $TR -m time_series_spark_b200.scorer_driver $D/scorer.yaml 2>&1 | grep -E "Error|error" | head
python tools/e2e_jobs.py check $D 2000 96
