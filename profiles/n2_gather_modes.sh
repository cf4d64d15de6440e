run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 100 --warmup 10 --no-e2e --no-dropin --no-cpu-baseline "${@:2}" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('${*:2}', round(d['ms_per_step'],4), '%.3e'%d['value'])"; }
run 29521
run 29522 --reserve-sms 2
run 29523 --reserve-sms 4
run 29524 --reserve-sms 8
run 29525 --gather serial
run 29526
