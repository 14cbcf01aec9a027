#!/bin/bash
set -u
TAG=${1:-exp}
mkdir -p gpurun_out
for v in default emu0 emu1 default emu0 emu1; do
  if [ $v = default ]; then unset OSD_LIB_PATH; else export OSD_LIB_PATH=$PWD/osu-dreamer_b200/ab/libosd_$v.so; fi
  timeout 200 python bench.py --steps 4 --warmup 3 --no-sampling --no-cpu > gpurun_out/${TAG}_bench_$v.json 2>> gpurun_out/${TAG}_bench.err
  python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench_$v.json'))
print('$v', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['roofline']['kernels'])
PY
done
