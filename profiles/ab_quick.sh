python -m pytest tests/test_gpu_parity.py -x -q -k "fused_vs_oracle or tensor_core_engine_token or config4_long or token_axis_edges or seeded" 2>&1 | tail -2
python profiles/config_sweep.py 2>&1 | tail -7
python bench.py --steps 60 --warmup 6 --no-cpu-baseline --no-configs --no-e2e --no-dropin | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'])"
python profiles/prior_tc_stats.py 2>&1 | grep -i "DP warp\|starved"
