# usage: tools/ncu_cfg.sh CFG KERNEL_REGEX OUTNAME [frames]   (run on the GPU box; leaves CSV exports in gpurun_out/)
export QB_CFG=$1 QB_FRAMES=${4:-48} QB_ITERS=2
R=gpurun_out/$3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -s 3 -o $R -f python tools/quick_bench.py > $R.log 2>&1; tail -1 $R.log
ncu -i $R.ncu-rep --page raw --csv > $R.raw.csv 2>/dev/null
ncu -i $R.ncu-rep --page source --csv --print-source sass > $R.src.csv 2>/dev/null
rm -f $R.ncu-rep
