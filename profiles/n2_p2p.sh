timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/multi_gpu_peer_gather.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -6
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 100 --warmup 10 --no-dropin --no-cpu-baseline "${@:2}" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('${*:2}', round(d['ms_per_step'],4), '%.3e'%d['value'], 'e2e', round(d['e2e']['ms_per_step'],3))"; }
run 29572 --gather p2p
run 29573 --gather serial
run 29574 --gather p2p
