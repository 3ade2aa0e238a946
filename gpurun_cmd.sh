#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(modwt_vs_oracle or fused_fir2d or fused_fir3d or fused_lift2d or fastpass_wpt) and fast" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.log
for spec in "lift2d --n 256 --batch 2 --levels 3" "filter2d --n 256 --batch 2 --levels 2 --wavelet db6" "wpt --n 8192 --batch 2" "modwt --n 32768 --batch 2 --levels 12" "filter3d --n 128 --levels 1"; do
  timeout 300 $CS --tool racecheck --error-exitcode 9 --print-limit 3 python tools/run_once.py --kind $spec > gpurun_out/race_tmp.log 2>&1; echo "racecheck [$spec] rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|max roundtrip" gpurun_out/race_tmp.log | tail -2; cat gpurun_out/race_tmp.log >> gpurun_out/sanitizer_racecheck.log
done
