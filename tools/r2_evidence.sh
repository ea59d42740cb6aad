# GPU box: round-2 evidence -- ncu launch list of the bench command, ncu --set full of the kernels DESIGN.md discusses
mkdir -p gpurun_out/r2ev; rm -f gpurun_out/r2ev/*
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2ev/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2ev/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
bash tools/ncu_cfg.sh c2 vpp_fused r2ev/c2 768
bash tools/ncu_cfg.sh c3 vpp_march r2ev/c3 32
bash tools/ncu_cfg.sh x:AREA:256x256:BGR24:PLANAR:f32 vpp_march r2ev/area256 160
bash tools/ncu_cfg.sh x:AREA:128x128:BGR24:PLANAR:f32 vpp_march r2ev/area128 160
bash tools/ncu_cfg.sh x:BILINEAR:3840x2160:RGB24:MERGED:u8 vpp_fused r2ev/bil2160 48
bash tools/ncu_cfg.sh x:BICUBIC:512x512:BGR24:PLANAR:f32 vpp_march r2ev/bic512 160
ls -la gpurun_out/r2ev
