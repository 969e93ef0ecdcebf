#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (small shapes): memcheck, then racecheck (shared-memory hazards).
# Output: gpurun_out/san_r2_{mem,race}.log
SEL='wavenet_tensor_core_vs_oracle_fp64 and (256-256-10 or 128-256-5) or tensor_core_mixing_vs_oracle_fp64 and 128-512 or other_sequence_lengths_vs_reference_golden and (tiny_unet-2304 or tiny_snet-160)'
SELOPS='cauchy_backward_vs_autograd_complex128 and 32-1000 or mel_front_end_vs_reference_golden and small or cauchy_shapes and 5-3001'
for tool in memcheck racecheck; do
  out=gpurun_out/san_r2_${tool:0:4}.log
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "$SEL" > $out 2>&1
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "$SELOPS" >> $out 2>&1
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard" $out | head -20
done
