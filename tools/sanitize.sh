#!/bin/bash
# compute-sanitizer over the tcgen05 / TMA kernels at small shapes: memcheck and racecheck of the
# attention forward (variant 7), the single-pass and two-kernel backward, the GEMM (store, split-K reduce-add, QKV epilogue)
# and one whole model forward + backward.  Logs -> gpurun_out/<tag>_sanitizer_<tool>_<what>.log ; each run is time-boxed.
TAG=${1:-r02}
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  for what in attn gemm model; do
    [ "$tool" != memcheck ] && [ "$what" = model ] && continue
    log=gpurun_out/${TAG}_sanitizer_${tool}_${what}.log
    timeout 300 compute-sanitizer --tool $tool --print-limit 30 --launch-timeout 120 python tools/sanitize_target.py $what > $log 2>&1
    echo "rc=$?" >> $log
    echo "== $tool $what: $(grep -c 'ERROR SUMMARY' $log) summary line(s): $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $log | tail -1) $(tail -1 $log)"
  done
done
