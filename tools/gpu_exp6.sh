#!/bin/bash
set -u
TAG=${1:-exp}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_trainer.py -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log | cut -c1-200
timeout 200 python bench.py --steps 4 --warmup 3 --no-sampling --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-330 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-sampling --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/agg_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_agg.txt; head -22 gpurun_out/${TAG}_launches_agg.txt
