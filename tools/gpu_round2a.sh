#!/bin/bash
# round-2 first GPU visit: new parity tests, full bench line (+ kernel-to-beat, cpu baseline), reference arm, CUPTI timeline,
# compute-sanitizer.  Everything lands in gpurun_out/<tag>_*
set -u
TAG=${1:-r02a}
mkdir -p gpurun_out
( nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,memory.total --format=csv; nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|Thread|Core" ) > gpurun_out/${TAG}_box.txt 2>&1
( time python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -25 gpurun_out/${TAG}_pytest.log | cut -c1-250
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-1200 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python tools/step_timeline.py 16 8192 3 > gpurun_out/${TAG}_step_timeline.json 2> gpurun_out/${TAG}_step_timeline.err
head -c 1500 gpurun_out/${TAG}_step_timeline.json; tail -3 gpurun_out/${TAG}_step_timeline.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cut -c1-800 gpurun_out/${TAG}_bench_reference.json
bash tools/sanitize.sh ${TAG}
ls -la gpurun_out | tail -20
