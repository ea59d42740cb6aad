# same-box A/B of the working tree's library against tensor_stream_b200/libtsvpp_prev.so (developer tool)
for c in "$@"; do
  for rep in 1 2; do
    echo "$c new : $(QB_CFG=$c timeout 120 python tools/quick_bench.py 2>&1 | tail -1)"
    echo "$c prev: $(QB_CFG=$c TSVPP_LIB=$PWD/tensor_stream_b200/libtsvpp_prev.so timeout 120 python tools/quick_bench.py 2>&1 | tail -1)"
  done
done
