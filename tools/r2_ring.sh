# GPU box: full GPU test-suite, then ring throughput in every mode (developer tool)
mkdir -p gpurun_out/r2d; rm -f gpurun_out/r2d/*
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2d/pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2d/rc.txt
grep -E 'passed|failed|FAILED|Error' gpurun_out/r2d/pytest_gpu.log | tail -20
B=tests/cpp/ring_main
for n in 1 4 8 16 32; do
  for mode in "TSVPP_ASYNC=0" "TSVPP_ASYNC=1" "TSVPP_COALESCE=0 TSVPP_ASYNC=0" "TSVPP_COALESCE=0 TSVPP_ASYNC=1"; do
    echo "$mode : $(env $mode TSVPP_COALESCE_STATS=1 timeout 120 $B run $n 4000 0 | tr '\n' ' ')" >> gpurun_out/r2d/ring.txt
  done
done
cat gpurun_out/r2d/ring.txt
for c in x:AREA:256x256:BGR24:PLANAR:f32 x:AREA:128x128:BGR24:PLANAR:f32 c3; do QB_CFG=$c timeout 120 python tools/quick_bench.py 2>&1 | tail -1; done
