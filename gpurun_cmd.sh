timeout 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=5 -x -k "lift" > gpurun_out/pytest.log 2>&1; echo pytest_rc=$?
tail -5 gpurun_out/pytest.log
timeout 300 python - <<'PY'
import sys, time, ctypes as C
sys.path.insert(0,'.')
import torch, numpy as np
import wavelets_b200 as wb
from wavelets_b200 import _lib
L=_lib.lib()
wl=wb.wavelet(wb.WT.cdf97, wb.WT.Lifting)
for dt,B in ((torch.float32,16),(torch.float64,8)):
    x=torch.randn((B,4096,4096),dtype=dt,device='cuda').permute(2,1,0)
    for _ in range(2):
        y=wb.dwtc(x,wl,8); xr=wb.idwtc(y,wl,8)
    torch.cuda.synchronize()
    L.wb200_profile_enable(1)
    for _ in range(5):
        y=wb.dwtc(x,wl,8); xr=wb.idwtc(y,wl,8)
    torch.cuda.synchronize(); L.wb200_profile_enable(0)
    buf=C.create_string_buffer(1<<14); nb=L.wb200_profile_collect(buf,len(buf))
    tot={}
    for ln in buf.raw[:nb].decode().splitlines():
        nm,c,ms=ln.split(); tot[nm]=(int(c),float(ms)/5)
    esz=x.element_size(); bytes_pass=2*esz*4096*4096*B
    fwd=sum(v[1] for k,v in tot.items() if 'fwd' in k or 'forward' in k or 'analysis' in k)
    inv=sum(v[1] for k,v in tot.items() if 'inv' in k or 'synthesis' in k)
    print(dt, tot)
    print('  fwd GB/s', bytes_pass/fwd/1e6, 'inv GB/s', bytes_pass/inv/1e6, 'rt', float((xr-x).abs().max()))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lift2d' -c 2 -f -o gpurun_out/r01_lift2d_f32 python tools/run_once.py --kind lift2d --dtype f32 --batch 4 > gpurun_out/ncu_2d.log 2>&1; echo ncu_rc=$?
