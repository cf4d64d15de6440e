for pf in 0 1; do
echo "== MAS_TC_L2_PREFETCH=$pf"
MAS_TC_L2_PREFETCH=$pf python bench.py --steps 120 --warmup 12 --no-cpu-baseline --no-configs --no-e2e --no-dropin | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench B=1024 ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'])"
MAS_TC_L2_PREFETCH=$pf python bench.py --batch 128 --steps 240 --warmup 24 --no-cpu-baseline --no-configs --no-e2e --no-dropin | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench B=128 shard ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'])"
done
