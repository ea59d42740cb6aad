mkdir -p gpurun_out/r2j; rm -f gpurun_out/r2j/*
timeout 600 python -m pytest tests/test_parity_gpu.py -q -k "yuv or uyvy" > gpurun_out/r2j/pytest.log 2>&1; echo "rc=$?"
grep -E 'passed|failed|FAILED|Error' gpurun_out/r2j/pytest.log | tail -20
for c in x:NEAREST:0x0:UYVY:MERGED:u8 x:NEAREST:0x0:YUV444:MERGED:u8 x:NEAREST:0x0:UYVY:MERGED:f32 x:NEAREST:0x0:YUV444:MERGED:f32; do
  for e in "" "TSVPP_YUV_GENERIC=1"; do echo "$e $(env $e QB_CFG=$c QB_FRAMES=160 timeout 120 python tools/quick_bench.py 2>&1 | tail -1)" >> gpurun_out/r2j/qb.txt; done
done
for c in x:BILINEAR:3840x2160:RGB24:MERGED:u8 x:BILINEAR:1280x720:RGB24:MERGED:u8 x:AREA:256x256:BGR24:PLANAR:f32 x:AREA:128x128:BGR24:PLANAR:f32; do
  for lib in "" _B; do echo "lib=$lib $(TSVPP_LIB=$PWD/tensor_stream_b200/libtsvpp$lib.so QB_CFG=$c QB_FRAMES=96 timeout 120 python tools/quick_bench.py 2>&1 | tail -1)" >> gpurun_out/r2j/qb.txt; done
done
cat gpurun_out/r2j/qb.txt
