#!/bin/bash
# compute-sanitizer over CI-size shapes of the hand-written tcgen05 / TMA kernels (K1 v2, K5, K4) and the dense kernel.
# usage (on a GPU box): bash tools/sanitize.sh <outdir>   -> <outdir>/sanitize_{memcheck,racecheck,synccheck}.log + summary
OUT=${1:-gpurun_out}
SEL='(test_conv2d_tcgen05_bit_exact and v2 and (n1_16x16 or n3_16x16_64 or n5_8x8 or n2_32x32_256-256_w8a8_pool)) or (test_first_layer_tcgen05_bit_exact and not n150) or (test_conv2d_f32_tcgen05_tolerance and (n2_32x32 or n5_8x8)) or (test_dense_head)'
for tool in memcheck racecheck synccheck; do
  timeout 480 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "$SEL" > $OUT/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" >> $OUT/sanitize_$tool.log
done
for tool in memcheck racecheck synccheck; do
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitize_$tool.log | tail -1) | $(grep -E 'passed|failed' $OUT/sanitize_$tool.log | tail -1) | $(tail -1 $OUT/sanitize_$tool.log)"
done | tee $OUT/sanitize_summary.txt
