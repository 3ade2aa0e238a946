timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=5 -x > gpurun_out/pytest.log 2>&1; echo pytest_rc=$?
tail -3 gpurun_out/pytest.log
echo "== persist=0"; timeout 200 python tools/bench2d.py 2>&1 | grep -E "GB/s"
echo "== persist=1"; WB200_LIFT2D_PERSIST=1 timeout 200 python tools/bench2d.py 2>&1 | grep -E "GB/s"
WB200_LIFT2D_PERSIST=1 timeout 600 python -m pytest tests -m gpu -q -x -k "fused_lift2d or full_size_2d" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f32.log 2> gpurun_out/bench.err; echo bench_rc=$?
timeout 400 python bench.py --steps 10 --warmup 3 --dtype f64 --no-extras > gpurun_out/bench_f64.log 2>> gpurun_out/bench.err
for dt in f32 f64; do python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$dt.log').read().strip().splitlines()[-1])
print('$dt value',round(d['value']),'pairGB/s',round(d['achieved_gbs_pair']),'frac',round(d['achieved_gbs_pair']/6570,3),d['roofline']['kernel'],round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']),'cpu',d['cpu_baseline'] and round(d['cpu_baseline']['value']),d['clocks'])
for k,v in d['extras'].items(): print('   ',k,round(v['msamples_per_s_pair']),round(v['frac_of_hbm_peak'],3))
PY
done
for dt in f32 f64; do
  B=512; [ $dt = f64 ] && B=256
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_ana_tiles|k_syn_tiles' -c 2 -f -o gpurun_out/r01b_fused1d_$dt python tools/run_once.py --kind filter1d --dtype $dt --batch $B > gpurun_out/ncu_$dt.log 2>&1; echo ncu_${dt}_rc=$?
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 3 --batch 1024 --no-extras > gpurun_out/bench_under_ncu.log 2>&1; echo launches_rc=$?
