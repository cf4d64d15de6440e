python -m pytest tests/test_gpu_parity.py tests/test_gpu_alignment.py -x -q -k "fused or tensor_core or cluster or config4 or config3 or long_text or engines or host_buffer" 2>&1 | tail -2
python bench.py --steps 60 --warmup 6 --no-cpu-baseline --no-e2e --no-dropin | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench B=1024 ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms']); print({k:(round(v['dropin_ms'],4),round(v['fused_ms'],4)) for k,v in d['configs'].items()})"
python profiles/prior_tc_stats.py 2>&1 | grep -i "DP warp 0\|starved of\|epilogue\|wait D\|ld + adds\|ring stage"
