timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=5 -x > gpurun_out/pytest.log 2>&1; echo pytest_rc=$?
tail -3 gpurun_out/pytest.log
echo "== 2-D occ5=1"; timeout 200 python tools/bench2d.py 2>&1 | grep -E "GB/s"
echo "== 2-D occ5=0"; WB200_LIFT2D_OCC5=0 timeout 200 python tools/bench2d.py 2>&1 | grep -E "GB/s"
echo "== 1-D f32 default (inv tile 16384/tail 16384)"; timeout 300 python tools/tune_fused1d.py --dtype f32 --batch 4096 2>&1 | head -0
timeout 400 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_f32.log 2>> gpurun_out/bench.err
WB200_TILE_F32_INV=8192 WB200_TAILMAX_F32_INV=32768 timeout 400 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_f32_b.log 2>> gpurun_out/bench.err
for f in bench_f32 bench_f32_b; do python - <<PY
import json
d=json.loads(open('gpurun_out/$f.log').read().strip().splitlines()[-1])
print('$f value',round(d['value']),'pairGB/s',round(d['achieved_gbs_pair']),'frac',round(d['achieved_gbs_pair']/6570,3),d['roofline']['all_kernels_ms'], d['clocks'])
PY
done
