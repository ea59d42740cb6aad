mkdir -p gpurun_out/r2n; rm -f gpurun_out/r2n/*
python -c "from tests.test_ring_gpu import _build; print(_build())" > /dev/null 2>&1
uptime > gpurun_out/r2n/host.txt
B=tests/cpp/ring_main
for rep in 1 2; do for linger in 4 0 8; do for src in ring direct; do for a in 1 0; do
    echo "rep=$rep linger=$linger src=$src async=$a : $(TSVPP_CO_LINGER_US=$linger RING_SOURCE=$src TSVPP_ASYNC=$a RING_OUT=pool TSVPP_COALESCE_STATS=1 timeout 60 $B run 8 30000 0 | tr '\n' ' ' | sed -e 's/published.*coalesce_stats//')" >> gpurun_out/r2n/ring.txt
done; done; done; done
echo "1 consumer: $(RING_SOURCE=direct TSVPP_ASYNC=0 RING_OUT=pool TSVPP_COALESCE_STATS=1 timeout 60 $B run 1 20000 0 | tr '\n' ' ' | sed -e 's/published.*coalesce_stats//')" >> gpurun_out/r2n/ring.txt
cat gpurun_out/r2n/host.txt gpurun_out/r2n/ring.txt
