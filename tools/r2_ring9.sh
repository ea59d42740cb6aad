mkdir -p gpurun_out/r2o; rm -f gpurun_out/r2o/*
python -c "from tests.test_ring_gpu import _build; print(_build())" > /dev/null 2>&1
B=tests/cpp/ring_main
for rep in 1 2; do for cfg in "1 4" "1 12" "1 20" "2 4" "3 4"; do set -- $cfg; for a in 1 0; do
    echo "rep=$rep depth=$1 linger=$2 async=$a : $(TSVPP_CO_DEPTH=$1 TSVPP_CO_LINGER_US=$2 TSVPP_ASYNC=$a RING_OUT=pool TSVPP_COALESCE_STATS=1 timeout 60 $B run 8 30000 0 | tr '\n' ' ' | sed -e 's/published.*coalesce_stats//')" >> gpurun_out/r2o/ring.txt
done; done; done
cat gpurun_out/r2o/ring.txt
