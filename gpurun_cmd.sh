timeout 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=10 -x > gpurun_out/pytest.log 2>&1; echo pytest_rc=$?
tail -4 gpurun_out/pytest.log
for dt in f32 f64; do
  B=512; [ $dt = f64 ] && B=256
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_ana_tiles|k_syn_tiles' -c 2 -f -o gpurun_out/r01_fused1d_$dt python tools/run_once.py --kind filter1d --dtype $dt --batch $B > gpurun_out/ncu_$dt.log 2>&1; echo ncu_${dt}_rc=$?
  tail -2 gpurun_out/ncu_$dt.log
done
ls -la gpurun_out/*.ncu-rep
