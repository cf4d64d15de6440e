# quick same-box check of a kernel change: parity of the fused paths, bench (one launch alone / pipelined), the
# 128-utterance shard of the batch-sharded step, configs 1-4
python -m pytest tests/test_gpu_parity.py tests/test_gpu_alignment.py -x -q -k "fused or tensor_core or cluster or config4 or config3 or long_text or engines or host_buffer or path_out or degenerate or peer" 2>&1 | tail -2
python bench.py --steps 60 --warmup 6 --no-cpu-baseline --no-e2e --no-dropin | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench B=1024 ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms']); print({k:(round(v['dropin_ms'],4),round(v['fused_ms'],4)) for k,v in d['configs'].items()})"
python bench.py --batch 128 --steps 240 --warmup 24 --no-cpu-baseline --no-configs --no-e2e --no-dropin | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench B=128 shard ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'])"
