timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=5 -x > gpurun_out/pytest.log 2>&1; echo pytest_rc=$?
tail -3 gpurun_out/pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f32.log 2> gpurun_out/bench.err; echo rc=$?
timeout 400 python bench.py --steps 10 --warmup 3 --dtype f64 --no-extras > gpurun_out/bench_f64.log 2>> gpurun_out/bench.err; echo rc=$?
for dt in f32 f64; do python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$dt.log').read().strip().splitlines()[-1])
print('$dt value',round(d['value']),'pairGB/s',round(d['achieved_gbs_pair']),'frac',round(d['achieved_gbs_pair']/6570,3),d['roofline']['kernel'],round(d['roofline']['frac'],3),d['roofline']['all_kernels_ms'],d['extras'], d['clocks'])
PY
done
