mkdir -p gpurun_out/r2s; rm -f gpurun_out/r2s/*
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 97 --print-limit 20 \
   python -m pytest tests/test_march_gpu.py tests/test_parity_gpu.py -q -x -k "integer_kernel_equals_oracle or every_output_format or unaligned_frames or packed_yuv_vector or (cuda_equals_oracle and (C2 or C3 or sweep_area_256 or crop_area or upscale))" \
   > gpurun_out/r2s/memcheck.log 2>&1; echo "memcheck rc=$?"
tail -15 gpurun_out/r2s/memcheck.log
