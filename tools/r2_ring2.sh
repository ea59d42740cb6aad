mkdir -p gpurun_out/r2e; rm -f gpurun_out/r2e/*
python -c "from tests.test_ring_gpu import _build; print(_build())" > /dev/null 2>&1
B=tests/cpp/ring_main
for n in 1 2 4 8 16; do
  for mode in "TSVPP_ASYNC=0" "TSVPP_ASYNC=1" "TSVPP_COALESCE=0 TSVPP_ASYNC=1"; do
    echo "$mode : $(env $mode RING_OUT=${RING_OUT:-pool} RING_FREE=async TSVPP_COALESCE_STATS=1 timeout 120 $B run $n 20000 0 | tr '\n' ' ')" >> gpurun_out/r2e/ring.txt
  done
done
echo "verify: $(TSVPP_ASYNC=1 RING_FREE=async RING_OUT=pool timeout 120 $B run 8 256 1 | tr '\n' ' ')" >> gpurun_out/r2e/ring.txt
cat gpurun_out/r2e/ring.txt
