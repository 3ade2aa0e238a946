#!/bin/bash
# compute-sanitizer over the round-2 kernels (small cases; each invocation bounded by its own timeout)
cd "$(dirname "$0")/.."
run() {  # tool, -k expression
  echo "== $1: $2"
  timeout 600 compute-sanitizer --tool $1 --launch-timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -k "$2" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error:|hazard" | tail -4
}
run memcheck "onepass_fir3d and db6 and float32 and strict"
run memcheck "fused_lift1d and cdf97 and float32 and strict and 256"
run memcheck "fused_lift2d_vs_oracle and cdf97 and float32 and strict and 256-8-3"
run memcheck "denoise_vs_oracle and float32 and hard and strict"
run racecheck "onepass_fir3d_chunks and 2"
run racecheck "fused_lift1d and cdf97 and float32 and strict and 256"
run racecheck "fused_lift2d_vs_oracle and cdf97 and float32 and strict and 256-8-3"
run racecheck "noisest_vs_oracle and float32 and strict"
