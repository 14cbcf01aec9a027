#!/bin/bash
# experiment visit: backward parity tests, short bench, GEMM shape timings, ncu speed-of-light of the non-attention kernels
set -u
TAG=${1:-exp}
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_backward.py tests/test_gpu_trainer.py -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --steps 4 --warmup 3 --no-sampling --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-400 gpurun_out/${TAG}_bench.json
python tools/gemm_ab.py > gpurun_out/${TAG}_gemm.txt 2>&1
cat gpurun_out/${TAG}_gemm.txt
timeout 600 ncu --profile-from-start off --section SpeedOfLight --metrics dram__bytes_read.sum,dram__bytes_write.sum \
    -k regex:'^(?!.*attn_)' --clock-control none --csv --log-file gpurun_out/${TAG}_sol.csv python tools/prof_step.py 16 8192 \
    > gpurun_out/${TAG}_sol.log 2>&1
tail -2 gpurun_out/${TAG}_sol.log
python tools/agg_ncu_csv.py gpurun_out/${TAG}_sol.csv 6532.5 gpurun_out/${TAG}_sol_summary.json | head -40
