timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=5 -x > gpurun_out/pytest.log 2>&1; echo pytest_rc=$?
tail -5 gpurun_out/pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_lift2d and strict and (256 or 128 or 384)" > gpurun_out/sanitizer.log 2>&1; echo memcheck_rc=$?
tail -3 gpurun_out/sanitizer.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lift2d_fwd_tma|k_lift2d_inv_tma' -c 2 -f -o gpurun_out/r01_lift2d_f32_v3 python tools/run_once.py --kind lift2d --dtype f32 --batch 4 > gpurun_out/ncu_2d.log 2>&1; echo ncu_rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lift2d_inv_tma' -c 1 -f -o gpurun_out/r01_lift2d_f32_v3_inv python tools/run_once.py --kind lift2d --dtype f32 --batch 4 > gpurun_out/ncu_2d.log 2>&1; echo ncu_rc=$?
