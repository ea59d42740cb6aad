mkdir -p gpurun_out/r2h; rm -f gpurun_out/r2h/*
python -c "from tests.test_ring_gpu import _build; print(_build())" > /dev/null 2>&1
B=tests/cpp/ring_main
for rep in 1 2 3; do for n in 4 8; do for a in 2 1 0; do
    echo "rep=$rep async=$a : $(TSVPP_ASYNC=$a RING_OUT=pool TSVPP_COALESCE_STATS=1 timeout 60 $B run $n 40000 0 | tr '\n' ' ' | sed -e 's/published.*coalesce_stats//')" >> gpurun_out/r2h/ring.txt
done; done; done
cat gpurun_out/r2h/ring.txt
timeout 200 python -m pytest tests/test_ring_gpu.py -q -x 2>&1 | tail -3
