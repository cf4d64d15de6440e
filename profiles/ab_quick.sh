python -m pytest tests/test_gpu_parity.py tests/test_gpu_alignment.py -x -q -k "fused or tensor_core or cluster or config4 or config3 or long_text or engines or host_buffer or graph or peer or alignment" 2>&1 | tail -2
for st in 0 1; do
echo "== MAS_TC_STAGGER=$st"
MAS_TC_STAGGER=$st python bench.py --steps 60 --warmup 6 --no-cpu-baseline --no-e2e --no-dropin | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench B=1024 ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms']); print({k:(round(v['dropin_ms'],4),round(v['fused_ms'],4)) for k,v in d['configs'].items()})"
MAS_TC_STAGGER=$st python bench.py --batch 128 --steps 240 --warmup 24 --no-cpu-baseline --no-configs --no-e2e --no-dropin | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench B=128 shard ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'])"
MAS_TC_STAGGER=$st python profiles/prior_tc_stats.py 2>&1 | grep -i "DP warp 0\|starved of\|MMA lane total\|issue + commit\|wait D\|ld + adds"
done
