mkdir -p gpurun_out/r2y; rm -f gpurun_out/r2y/*
timeout 300 python -m pytest tests/test_parity_gpu.py -q -k "two_contexts" > gpurun_out/r2y/pytest_2gpu.log 2>&1; echo "2gpu test rc=$?"; tail -2 gpurun_out/r2y/pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2y/bench_ours_2gpu.json 2> gpurun_out/r2y/bench_ours_2gpu.err; echo "bench2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2y/bench_ref_2gpu.json 2> gpurun_out/r2y/bench_ref_2gpu.err; echo "ref2 rc=$?"
tail -c 300 gpurun_out/r2y/bench_ours_2gpu.err
