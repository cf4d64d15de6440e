# same-box A/B of builds of the fused kernel: bash profiles/ab_variants.sh "A D" (libmas_ab_<v>.so, D = the product)
for rep in 1 2; do
for v in ${1:-A D}; do
  case $v in
    D) unset MAS_LIB_PATH; export MAS_PRIOR_TC2=0;;
    E) unset MAS_LIB_PATH; export MAS_PRIOR_TC2=1;;
    *) export MAS_LIB_PATH=art_tts_b200/lib/libmas_ab_$v.so MAS_PRIOR_TC2=0;;
  esac
  python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-dropin 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', $rep, round(d['ms_per_step'],4))"
done; done
