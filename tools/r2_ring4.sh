mkdir -p gpurun_out/r2g; rm -f gpurun_out/r2g/*
python -c "from tests.test_ring_gpu import _build; print(_build())" > /dev/null 2>&1
(echo "nproc $(nproc)"; echo "cpu.max $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; echo "cpuset $(cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null)"; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))"; uptime; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|Thread") > gpurun_out/r2g/host.txt 2>&1
cat gpurun_out/r2g/host.txt
B=tests/cpp/ring_main
for rep in 1 2 3; do
for n in 8 16; do
    echo "rep=$rep depth=1 async : $(TSVPP_ASYNC=1 TSVPP_CO_DEPTH=1 RING_OUT=pool TSVPP_COALESCE_STATS=1 timeout 120 $B run $n 20000 0 | tr '\n' ' ' | sed -e 's/published.*coalesce_stats//')" >> gpurun_out/r2g/ring.txt
    echo "rep=$rep depth=1 sync  : $(TSVPP_ASYNC=0 TSVPP_CO_DEPTH=1 RING_OUT=pool TSVPP_COALESCE_STATS=1 timeout 120 $B run $n 20000 0 | tr '\n' ' ' | sed -e 's/published.*coalesce_stats//')" >> gpurun_out/r2g/ring.txt
done
done
cat gpurun_out/r2g/ring.txt
