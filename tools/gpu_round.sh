#!/bin/bash
# one GPU-box visit: parity tests, bench line (both arms), ncu launch list of the bench command, ncu --set full of attention,
# seq-len sweep (BASELINE configs[4]), predict latency; everything lands in gpurun_out/<tag>_*
set -u
TAG=${1:-r02z}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log | cut -c1-200
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-1500 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cut -c1-600 gpurun_out/${TAG}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-sampling --no-cpu --no-alt --no-kernel-to-beat > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/agg_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_agg.txt; head -12 gpurun_out/${TAG}_launches_agg.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ --launch-skip 7 -c 7 -f -o gpurun_out/${TAG}_attn \
    python tools/prof_attn.py 16 8192 > gpurun_out/${TAG}_ncu_attn.log 2>&1
timeout 600 python tools/seq_sweep.py > gpurun_out/${TAG}_seq_sweep.log 2>&1; cp gpurun_out/seq_sweep.json gpurun_out/${TAG}_seq_sweep.json
tail -4 gpurun_out/${TAG}_seq_sweep.log | cut -c1-300
python tools/predict_latency.py 4 4 8 2>&1 | tail -1; cp gpurun_out/predict_latency.json gpurun_out/${TAG}_predict_latency_8steps.json
python tools/predict_latency.py 4 4 64 2>&1 | tail -1; cp gpurun_out/predict_latency.json gpurun_out/${TAG}_predict_latency_64steps.json
ls -la gpurun_out | tail -14
# sanitizers over the kernels added late in the round (time-boxed)
for tool in memcheck racecheck; do
  log=gpurun_out/${TAG}_sanitizer_${tool}_new.log
  OSD_GEMM_PAIR=1 timeout 400 compute-sanitizer --tool $tool --print-limit 30 --launch-timeout 120 python tools/sanitize_target.py new > $log 2>&1
  echo "rc=$?" >> $log
  echo "== $tool new: $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $log | tail -1) $(tail -1 $log)"
done
