for cfg in "4 6" "6 8" "8 10" "8 12" "12 16" "16 16"; do
  set -- $cfg
  python bench.py --batch 128 --utt-per-cta $1 --streams $2 --steps 240 --warmup 24 --no-cpu-baseline --no-configs --no-e2e --no-dropin 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('upc $1 streams $2', round(d['ms_per_step'],5))"
done
