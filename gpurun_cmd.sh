#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "threshold or noisest or denoise" 2>&1 | tail -25 > gpurun_out/pytest_part.log; cat gpurun_out/pytest_part.log
