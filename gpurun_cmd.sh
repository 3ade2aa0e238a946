#!/bin/bash
# scratch: GPU validation of MODWT + the reworked 2-D level kernels
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "modwt or lift2d or full_size_2d or golden or lifting" 2>&1 | tail -15 > gpurun_out/pytest_part.log
cat gpurun_out/pytest_part.log
timeout 300 python tools/bench2d.py 16 8 --sweep > gpurun_out/bench2d.log 2>&1; cat gpurun_out/bench2d.log
timeout 300 python tools/bench2d.py 64 0 > gpurun_out/bench2d_b64.log 2>&1; cat gpurun_out/bench2d_b64.log
timeout 300 python tools/bench_modwt.py > gpurun_out/bench_modwt.log 2>&1; cat gpurun_out/bench_modwt.log
