# zero-copy end-to-end rate under different plan knobs (developer tool; run on the GPU box)
for knobs in "" "TSVPP_STAGES=3" "TSVPP_STAGES=4" "TSVPP_RP=1" "TSVPP_RP=1 TSVPP_STAGES=4" "TSVPP_TW=128" "TSVPP_CTAS_PER_SM=2" "TSVPP_CTAS_PER_SM=3"; do
  echo "[$knobs] $(env $knobs E2E_QUICK=1 python tools/e2e_variants.py 2>&1 | head -1)"
done
