timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=5 -x > gpurun_out/pytest.log 2>&1; echo pytest_rc=$?
tail -8 gpurun_out/pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(fastpass_wpt and strict) or (test_wpt_vs_oracle and strict)" > gpurun_out/sanitizer.log 2>&1; echo memcheck_rc=$?
tail -3 gpurun_out/sanitizer.log
timeout 300 python tools/bench_nd.py
