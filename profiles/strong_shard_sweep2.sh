# 128-utterance shard of the batch-sharded config 5 on one GPU: utterances per CTA x streams
for cfg in "3 6" "4 6" "4 8" "6 8" "8 10" "8 12"; do
  set -- $cfg
  python bench.py --batch 128 --utt-per-cta $1 --streams $2 --steps 240 --warmup 24 --no-cpu-baseline --no-configs --no-e2e --no-dropin 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('upc $1 streams $2', round(d['ms_per_step'],5))"
done
for cfg in "2 6" "3 6" "4 6"; do
  set -- $cfg
  python bench.py --batch 256 --utt-per-cta $1 --streams $2 --steps 240 --warmup 24 --no-cpu-baseline --no-configs --no-e2e --no-dropin 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=256 upc $1 streams $2', round(d['ms_per_step'],5))"
done
for cfg in "1 6" "2 6" "4 6"; do
  set -- $cfg
  python bench.py --batch 512 --utt-per-cta $1 --streams $2 --steps 240 --warmup 24 --no-cpu-baseline --no-configs --no-e2e --no-dropin 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=512 upc $1 streams $2', round(d['ms_per_step'],5))"
done
