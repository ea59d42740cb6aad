mkdir -p gpurun_out/r2m; rm -f gpurun_out/r2m/*
python -c "from tests.test_ring_gpu import _build; print(_build())" > /dev/null 2>&1
uptime > gpurun_out/r2m/host.txt
B=tests/cpp/ring_main
for rep in 1 2; do for src in ring direct; do for a in 2 1 0; do
    echo "rep=$rep src=$src async=$a : $(RING_SOURCE=$src TSVPP_ASYNC=$a RING_OUT=pool TSVPP_COALESCE_STATS=1 timeout 60 $B run 8 40000 0 | tr '\n' ' ' | sed -e 's/published.*coalesce_stats//')" >> gpurun_out/r2m/ring.txt
done; done; done
cat gpurun_out/r2m/host.txt gpurun_out/r2m/ring.txt
timeout 200 python -m pytest tests/test_ring_gpu.py -q -x 2>&1 | tail -3
