python -m pytest tests/test_gpu_parity.py -x -q -k "fused or tensor_core or cluster or config4" 2>&1 | tail -2
python profiles/prior_tc_stats.py 128 2>&1 | grep -i "mover\|global loads\|wait A\|split"
python profiles/cfg4_time.py 2>&1 | grep -i "cluster"
python profiles/config_sweep.py 2>&1 | tail -7
python bench.py --steps 60 --warmup 6 --no-cpu-baseline --no-configs --no-e2e --no-dropin | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'])"
python bench.py --batch 128 --steps 240 --warmup 24 --no-cpu-baseline --no-configs --no-e2e --no-dropin | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench B=128 shard ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'])"
