export QB_CFG=noresize QB_FRAMES=96 QB_ITERS=10
for np in 1 2; do for c in 0 4 5 6; do echo "NPROD=$np CTAS=$c: $(TSVPP_NPROD=$np TSVPP_CTAS_PER_SM=$c timeout 100 python tools/quick_bench.py 2>&1 | tail -2 | tr '\n' ' ' | sed -E 's/.*(threads=[0-9]+ ctas_per_sm=[0-9]+).*(ms .*)/\1 \2/')"; done; done
for np in 1 2; do echo "NPROD=$np RP=1 TW=256: $(TSVPP_NPROD=$np TSVPP_RP=1 timeout 100 python tools/quick_bench.py 2>&1 | tail -2 | tr '\n' ' ' | sed -E 's/.*(threads=[0-9]+ ctas_per_sm=[0-9]+).*(ms .*)/\1 \2/')"; done
TSVPP_NPROD=2 QB_FRAMES=48 QB_ITERS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:vpp_fused -c 1 -s 3 -o gpurun_out/nr_v7 -f python tools/quick_bench.py > gpurun_out/nr_ncu.log 2>&1; tail -2 gpurun_out/nr_ncu.log
