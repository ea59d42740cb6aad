mkdir -p gpurun_out/r2f; rm -f gpurun_out/r2f/*
python -c "from tests.test_ring_gpu import _build; print(_build())" > /dev/null 2>&1
B=tests/cpp/ring_main
for depth in 1 2 4; do
for n in 2 8 16; do
  for mode in "TSVPP_ASYNC=0" "TSVPP_ASYNC=1"; do
    echo "depth=$depth $mode : $(env $mode TSVPP_CO_DEPTH=$depth RING_OUT=pool TSVPP_COALESCE_STATS=1 timeout 120 $B run $n 20000 0 | tr '\n' ' ' | sed -e 's/published.*coalesce_stats//')" >> gpurun_out/r2f/ring.txt
  done
done
done
cat gpurun_out/r2f/ring.txt
