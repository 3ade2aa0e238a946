#!/bin/bash
# compute-sanitizer over the kernels touched last in round 2 (small cases; each invocation bounded by its own timeout):
# branch-free 1-D analysis level / four-pair synthesis, packet subtree kernels (shift-mask path, leaf 16x16 stage),
# fused K-level packet tile kernels
cd "$(dirname "$0")/.."
run() {  # tool, -k expression
  echo "== $1: $2"
  timeout 900 compute-sanitizer --tool $1 --launch-timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -k "$2" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error:|hazard" | tail -4
}
run memcheck "fused_tiles_vs_oracle and float32 and db4"
run memcheck "wpt_fused_levels and 16384-2-db2 and float32"
run memcheck "wpt_fused_levels and 32768-3-db4 and float64 and fast"
run memcheck "fastpass_wpt_full_tree and 4096-3-sym8 and float32"
run racecheck "fused_tiles_vs_oracle and float32 and db4 and fast"
run racecheck "wpt_fused_levels and 16384-2-db2 and float32 and fast"
run racecheck "fastpass_wpt_full_tree and 4096-3-sym8 and float32 and fast"
