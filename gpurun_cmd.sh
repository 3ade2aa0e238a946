timeout 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=5 -x -k "lift" > gpurun_out/pytest.log 2>&1; echo pytest_rc=$?
tail -5 gpurun_out/pytest.log
timeout 300 python tools/bench2d.py
