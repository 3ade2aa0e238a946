#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 300 python tools/bench_nd.py 2>&1 | head -1
timeout 300 python tools/bench2d.py 64 0 2>&1 | grep -v "^torch" 
timeout 300 python tools/bench2d.py 16 8 2>&1 | grep "whole calls"
timeout 300 python tools/bench_modwt.py 2>&1 | head -1
