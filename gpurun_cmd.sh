#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f32.log 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/bench_f32.log').read().strip().splitlines()[-1])
print('value',round(d['value'],1), 'frac',(d.get('roofline') or {}).get('frac'), 'e2e',(d.get('e2e') or {}).get('value'), 'clocks', d.get('clocks'))
for k,v in (d.get('extras') or {}).items(): print('   ',k, round(v['ms_per_pair'],3),'ms', round(v['achieved_gbs_pair'],1),'GB/s', round(v['frac_of_hbm_peak'],4))
PY
