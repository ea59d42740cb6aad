mkdir -p gpurun_out/r2l; rm -f gpurun_out/r2l/*
python -c "from tests.test_ring_gpu import _build; print(_build())" > /dev/null 2>&1
uptime > gpurun_out/r2l/host.txt
B=tests/cpp/ring_main
for rep in 1 2 3; do for n in 8; do for a in 2 1 0; do
    echo "rep=$rep async=$a : $(TSVPP_ASYNC=$a RING_OUT=pool TSVPP_COALESCE_STATS=1 timeout 60 $B run $n 40000 0 | tr '\n' ' ' | sed -e 's/published.*coalesce_stats//')" >> gpurun_out/r2l/ring.txt
done; done; done
echo "verify posted: $(TSVPP_ASYNC=2 RING_OUT=pool timeout 60 $B run 8 512 1 | tr '\n' ' ')" >> gpurun_out/r2l/ring.txt
cat gpurun_out/r2l/host.txt gpurun_out/r2l/ring.txt
