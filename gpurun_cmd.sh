#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f32.log 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; tail -c 300 gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_2gpu.log 2> gpurun_out/bench_ref_2gpu.err; tail -c 200 gpurun_out/bench_ref_2gpu.err
python - <<'PY'
import json
for f in ('bench_f32','bench_2gpu','bench_ref_2gpu'):
    try:
        d=json.loads([l for l in open(f'/root/repo/gpurun_out/{f}.log').read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'n_gpus', d.get('n_gpus'), 'value',round(d['value'],1), 'frac',(d.get('roofline') or {}).get('frac'), 'e2e',(d.get('e2e') or {}).get('value'), 'clocks', d.get('clocks'))
        for k,v in (d.get('extras') or {}).items(): print('   ',k, round(v['ms_per_pair'],3),'ms', round(v['achieved_gbs_pair'],1),'GB/s', round(v['frac_of_hbm_peak'],4))
    except Exception as e: print(f, 'ERR', e)
PY
