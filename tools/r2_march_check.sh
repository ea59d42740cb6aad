# GPU box: march-kernel parity + quick timings (developer tool)
mkdir -p gpurun_out/r2b; rm -f gpurun_out/r2b/qb.txt
timeout 600 python -m pytest tests/test_march_gpu.py -q > gpurun_out/r2b/pytest_march.log 2>&1; echo "march rc=$?" > gpurun_out/r2b/rc.txt
grep -E 'passed|failed|FAILED|Error' gpurun_out/r2b/pytest_march.log | tail -40
for c in ${QB_LIST:-c3 x:BICUBIC:512x512:BGR24:PLANAR:f32 x:AREA:256x256:BGR24:PLANAR:f32 x:AREA:128x128:BGR24:PLANAR:f32 x:BICUBIC:3840x2160:RGB24:MERGED:u8}; do
  QB_CFG=$c timeout 120 python tools/quick_bench.py 2>&1 | tail -2 >> gpurun_out/r2b/qb.txt
  QB_CFG=$c TSVPP_NO_MARCH=1 timeout 120 python tools/quick_bench.py 2>&1 | tail -1 >> gpurun_out/r2b/qb.txt
done
cat gpurun_out/r2b/qb.txt
if [ -n "$NCU_CFG" ]; then bash tools/ncu_cfg.sh $NCU_CFG vpp_march r2b/ncu_march ${NCU_FRAMES:-24}; fi
