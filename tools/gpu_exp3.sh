#!/bin/bash
set -u
TAG=${1:-exp}
mkdir -p gpurun_out
timeout 120 python tools/attn_fwd_ab.py > gpurun_out/${TAG}_attn_ab.txt 2>&1
cat gpurun_out/${TAG}_attn_ab.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-sampling --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/agg_launches.py gpurun_out/${TAG}_launches.csv | head -24
