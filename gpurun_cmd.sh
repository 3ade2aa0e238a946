#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "modwt" 2>&1 | tail -8 > gpurun_out/pytest_part.log
cat gpurun_out/pytest_part.log
timeout 300 python tools/bench_modwt.py > gpurun_out/bench_modwt.log 2>&1; cat gpurun_out/bench_modwt.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:k_lift2d_.*_tma -c 2 -o gpurun_out/r01c_lift2d_f32 python tools/run_once.py --kind lift2d --batch 8 --levels 1 > gpurun_out/ncu_a.log 2>&1; tail -2 gpurun_out/ncu_a.log
timeout 300 $NCU -k regex:k_lift2d_.*_tma -c 2 -o gpurun_out/r01c_fir2d_db4_f32 python tools/run_once.py --kind filter2d --batch 8 --levels 1 > gpurun_out/ncu_b.log 2>&1; tail -2 gpurun_out/ncu_b.log
timeout 300 $NCU -k regex:k_wpt_sub -c 2 -o gpurun_out/r01c_wpt_sub_f32 python tools/run_once.py --kind wpt --batch 512 > gpurun_out/ncu_c.log 2>&1; tail -2 gpurun_out/ncu_c.log
timeout 300 $NCU -k regex:modwt_group -c 6 -o gpurun_out/r01c_modwt_f32 python tools/run_once.py --kind modwt --batch 16 > gpurun_out/ncu_d.log 2>&1; tail -2 gpurun_out/ncu_d.log
ls -la gpurun_out/*.ncu-rep
