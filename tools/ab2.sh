for rep in 1 2 3; do for lib in "" _B; do for c in c3 x:BICUBIC:512x512:BGR24:PLANAR:f32 x:BICUBIC:3840x2160:RGB24:MERGED:u8; do
echo "lib=$lib $c $(TSVPP_LIB=$PWD/tensor_stream_b200/libtsvpp$lib.so QB_CFG=$c timeout 120 python tools/quick_bench.py 2>&1 | tail -1)"; done; done; done
