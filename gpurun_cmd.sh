#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "wpt" 2>&1 | tail -8 > gpurun_out/pytest_part.log
cat gpurun_out/pytest_part.log
timeout 300 python tools/bench_nd.py > gpurun_out/bench_nd.log 2>&1; cat gpurun_out/bench_nd.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_f32.log 2> gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/bench_f32.log').read().strip().splitlines()[-1])
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
for k,v in d.get('extras',{}).items(): print(k, round(v['ms_per_pair'],3),'ms', round(v['achieved_gbs_pair'],1),'GB/s', round(v['frac_of_hbm_peak'],4))
PY
tail -5 gpurun_out/bench.err
