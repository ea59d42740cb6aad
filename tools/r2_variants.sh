# GPU box: march tests on the default build, then quick_bench of A/B library variants (developer tool)
mkdir -p gpurun_out/r2c; rm -f gpurun_out/r2c/*
timeout 600 python -m pytest tests/test_march_gpu.py -q > gpurun_out/r2c/pytest_march.log 2>&1; echo "march rc=$?" > gpurun_out/r2c/rc.txt
grep -E 'passed|failed|FAILED|Error' gpurun_out/r2c/pytest_march.log | tail -20
for lib in "" _B _C; do
  for c in c3 x:BICUBIC:512x512:BGR24:PLANAR:f32 x:BICUBIC:3840x2160:RGB24:MERGED:u8; do
    for env in "" "TSVPP_MARCH_TW=128" "TSVPP_MARCH_RPG=4"; do
      echo "lib=$lib env=$env $(env $env QB_CFG=$c TSVPP_LIB=$PWD/tensor_stream_b200/libtsvpp$lib.so timeout 120 python tools/quick_bench.py 2>&1 | tail -2 | sed -e 's/.*kernel=march//' -e 's/ntiles.*slots/slots/' -e 's/qx=.*//' | tr '\n' ' ')" >> gpurun_out/r2c/qb.txt
    done
  done
done
cat gpurun_out/r2c/qb.txt
