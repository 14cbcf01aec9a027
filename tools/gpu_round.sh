#!/bin/bash
# one GPU-box visit: parity tests, bench line, ncu launch list of the bench command, ncu --set full of attention,
# seq-len sweep (BASELINE configs[4]); everything lands in gpurun_out/<tag>_*
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log | cut -c1-200
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json | cut -c1-1500
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-sampling --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/agg_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_agg.txt; head -12 gpurun_out/${TAG}_launches_agg.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ --launch-skip 7 -c 7 -f -o gpurun_out/${TAG}_attn \
    python tools/prof_attn.py 16 8192 > gpurun_out/${TAG}_ncu_attn.log 2>&1
timeout 600 python tools/seq_sweep.py > gpurun_out/${TAG}_seq_sweep.log 2>&1; cp gpurun_out/seq_sweep.json gpurun_out/${TAG}_seq_sweep.json
tail -4 gpurun_out/${TAG}_seq_sweep.log | cut -c1-300
ls -la gpurun_out | tail -12
