timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=5 -x -k "lift or golden" > gpurun_out/pytest.log 2>&1; echo pytest_rc=$?
tail -6 gpurun_out/pytest.log
timeout 300 python tools/bench2d.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lift2d_fwd|k_lift2d_inv' -c 2 -f -o gpurun_out/r01_lift2d_f32_v2 python tools/run_once.py --kind lift2d --dtype f32 --batch 4 > gpurun_out/ncu_2d.log 2>&1; echo ncu_rc=$?
