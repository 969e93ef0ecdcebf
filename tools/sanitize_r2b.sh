#!/bin/bash
# compute-sanitizer over the kernels (re)written in the second half of round 2: the persistent mixing kernel with TMEM operands
# and TMA-staged inputs (H = 64, 128), the double-buffered H = 256 kernel, the tcgen05 pools, the S4 convolution with its
# parked rows in tensor memory.  Output: gpurun_out/san_r2b_{memc,race}.log
SEL='baseline_size_vs_reference_golden and unet_d64 or tensor_core_mixing_vs_oracle_fp64 and (64-192 or 64-1024)'
for tool in memcheck racecheck; do
  out=gpurun_out/san_r2b_${tool:0:4}.log
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "$SEL" > $out 2>&1
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard" $out | head -20
done
