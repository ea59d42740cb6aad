mkdir -p gpurun_out/r2zz; rm -f gpurun_out/r2zz/*
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2zz/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2zz/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2zz/smoke.log 2>&1; echo "smoke rc=$?"
timeout 200 python tools/fuzz_parity.py --seconds 25 --seed 6 --profile wide-area > gpurun_out/r2zz/fuzz6.log 2>&1; echo "fuzz wide rc=$?"; tail -5 gpurun_out/r2zz/fuzz6.log | head -2
timeout 300 python tools/fuzz_parity.py --seconds 60 --seed 7 --profile broad --api all > gpurun_out/r2zz/fuzz7.log 2>&1; echo "fuzz api rc=$?"; grep -c MISMATCH gpurun_out/r2zz/fuzz7.log; grep "^cases" gpurun_out/r2zz/fuzz7.log
timeout 300 python tools/fuzz_parity.py --seconds 40 --seed 8 --profile edge --api all > gpurun_out/r2zz/fuzz8.log 2>&1; echo "fuzz api edge rc=$?"; grep -c MISMATCH gpurun_out/r2zz/fuzz8.log; grep "^cases" gpurun_out/r2zz/fuzz8.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2zz/bench_ref.json 2> gpurun_out/r2zz/bench_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2zz/bench_ours.json 2> gpurun_out/r2zz/bench_ours.err; echo "ours rc=$?"
