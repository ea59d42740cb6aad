# GPU box: the round's bench lines (both arms), smoke, and the ncu launch list of the bench command
mkdir -p gpurun_out/r2z; rm -f gpurun_out/r2z/*
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2z/bench_ref.json 2> gpurun_out/r2z/bench_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2z/bench_ours.json 2> gpurun_out/r2z/bench_ours.err; echo "ours rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'vpp_|nv12_' -c 400 --csv --log-file gpurun_out/r2z/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2z/bench_under_ncu.log 2>&1; echo "ncu rc=$?"
tail -c 600 gpurun_out/r2z/bench_ours.err
