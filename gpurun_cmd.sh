#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "threshold or noisest or denoise" 2>&1 | tail -5 > gpurun_out/pytest_part.log; cat gpurun_out/pytest_part.log
python - <<'PY'
import sys, time; sys.path.insert(0,'/root/repo')
import torch, wavelets_b200 as wb
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/reps
x=torch.randn(1<<24,device='cuda'); print('1-D 2^24 noTI', timed(lambda: wb.denoise(x)), 'ms;  TI 8 spins', timed(lambda: wb.denoise(x,TI=True)),'ms')
xi=torch.randn((1024,1024),device='cuda'); print('2-D 1024^2 TI 64 spins', timed(lambda: wb.denoise(xi,TI=True)),'ms; noTI', timed(lambda: wb.denoise(xi)),'ms')
import os; os.environ['WB200_DENOISE_CHUNK_MB']='0'; print('  one spin per batch:', timed(lambda: wb.denoise(xi,TI=True)),'ms')
PY
