mkdir -p gpurun_out/r2t; rm -f gpurun_out/r2t/*
for spec in "x:AREA:128x128:BGR24:PLANAR:f32|TSVPP_MARCH_TW=64|TSVPP_MARCH_TW=64 TSVPP_MARCH_G=1|TSVPP_MARCH_TW=32|TSVPP_MARCH_STAGES=3 TSVPP_MARCH_TW=64" "x:AREA:256x256:BGR24:PLANAR:f32|TSVPP_MARCH_TW=128|TSVPP_MARCH_G=4|TSVPP_MARCH_RPG=2|TSVPP_MARCH_STAGES=3" "x:BICUBIC:512x512:BGR24:PLANAR:f32|TSVPP_MARCH_RPG=4|TSVPP_MARCH_RPG=8|TSVPP_MARCH_G=1 TSVPP_MARCH_RPG=4" "x:NEAREST:1280x720:RGB24:MERGED:u8|TSVPP_BW=2|TSVPP_TW=128"; do
  cfg=${spec%%|*}; rest=${spec#*|}
  echo "base $(QB_CFG=$cfg timeout 120 python tools/quick_bench.py 2>&1 | tail -1)" >> gpurun_out/r2t/tune.txt
  IFS='|' read -ra envs <<< "$rest"
  for e in "${envs[@]}"; do
    echo "$e $(env $e QB_CFG=$cfg timeout 120 python tools/quick_bench.py 2>&1 | tail -2 | sed -e 's/.*kernel=march//' -e 's/out_kind=[0-9]* fill=bulk bw=[0-9]//' -e 's/ntiles.*slots/slots/' -e 's/qx=.*//' | tr '\n' ' ')" >> gpurun_out/r2t/tune.txt
  done
done
cat gpurun_out/r2t/tune.txt
