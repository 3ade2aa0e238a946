timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=5 -x > gpurun_out/pytest.log 2>&1; echo pytest_rc=$?
tail -4 gpurun_out/pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f32.log 2> gpurun_out/bench.err; echo bench_rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_f32.log').read().strip().splitlines()[-1])
print('f32 value',d['value'],'pairGB/s',d['achieved_gbs_pair'],'roof',d['roofline']['kernel'],d['roofline']['frac'],d['roofline']['all_kernels_ms'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'],d['extras'],d['clocks'])
PY
timeout 300 python bench.py --steps 10 --warmup 3 --dtype f64 --no-extras > gpurun_out/bench_f64.log 2>> gpurun_out/bench.err; echo bench64_rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_f64.log').read().strip().splitlines()[-1])
print('f64 value',d['value'],'pairGB/s',d['achieved_gbs_pair'],'roof',d['roofline']['kernel'],d['roofline']['frac'],d['roofline']['all_kernels_ms'],'e2e',d['e2e']['value'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --batch 1024 --no-extras > gpurun_out/bench_under_ncu.log 2>&1; echo launches_rc=$?
tail -3 gpurun_out/bench.err
