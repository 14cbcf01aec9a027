#!/bin/bash
# experiment visit: attention forward A/B (variants 4 / 6), backward parity tests, short bench
set -u
TAG=${1:-exp}
mkdir -p gpurun_out
timeout 120 python tools/attn_fwd_ab.py > gpurun_out/${TAG}_attn_ab.txt 2>&1
cat gpurun_out/${TAG}_attn_ab.txt
( time timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
timeout 200 python bench.py --steps 4 --warmup 3 --no-sampling --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-330 gpurun_out/${TAG}_bench.json
