mkdir -p gpurun_out/r2s
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 10 \
   python -m pytest tests/test_march_gpu.py tests/test_parity_gpu.py -q -x -k "(integer_kernel_equals_oracle and (bicubic_2x_up or area_2x or bilinear_2x_up)) or (unaligned_frames and bilinear) or (packed_yuv_vector and yuv444 and 64x16)" \
   > gpurun_out/r2s/racecheck.log 2>&1; echo "racecheck rc=$?"
tail -25 gpurun_out/r2s/racecheck.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 10 \
   python -m pytest tests/test_march_gpu.py tests/test_parity_gpu.py -q -x -k "(integer_kernel_equals_oracle and (bicubic_2x_up or area_2x or bilinear_2x_up)) or (unaligned_frames and bilinear) or (packed_yuv_vector and yuv444 and 64x16)" \
   > gpurun_out/r2s/synccheck.log 2>&1; echo "synccheck rc=$?"
tail -6 gpurun_out/r2s/synccheck.log
