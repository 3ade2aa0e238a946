#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest.log
cat gpurun_out/pytest.log
timeout 300 python tools/bench_modwt.py > gpurun_out/bench_modwt.log 2>&1; cat gpurun_out/bench_modwt.log
timeout 300 python tools/bench_nd.py > gpurun_out/bench_nd.log 2>&1; cat gpurun_out/bench_nd.log
