mkdir -p gpurun_out/r2i; rm -f gpurun_out/r2i/*
timeout 600 python -m pytest tests/test_march_gpu.py -q > gpurun_out/r2i/pytest_march.log 2>&1; echo "march rc=$?"
grep -E 'passed|failed|FAILED|Error' gpurun_out/r2i/pytest_march.log | tail -30
for c in x:BILINEAR:3840x2160:RGB24:MERGED:u8 x:BILINEAR:3840x2160:BGR24:PLANAR:f32 x:BILINEAR:1280x720:RGB24:MERGED:u8 x:BILINEAR:2560x1440:RGB24:MERGED:u8; do
  for e in "" "TSVPP_MARCH_BILINEAR=0" "TSVPP_MARCH_RPG=4" "TSVPP_MARCH_TW=128"; do
    echo "$e $(env $e QB_CFG=$c QB_FRAMES=96 timeout 120 python tools/quick_bench.py 2>&1 | tail -2 | sed -e 's/out_kind.*strip_rows/strip_rows/' -e 's/qx=.*//' -e 's/ntiles.*fetch/fetch/' | tr '\n' ' ')" >> gpurun_out/r2i/qb.txt
  done
done
cat gpurun_out/r2i/qb.txt
