nvidia-smi -L | wc -l; free -g | head -2 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 30 --warmup 5 --no-e2e 2>&1 | tail -1 | cut -c1-2500
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 3 --warmup 3 --e2e-steps 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('e2e', d['e2e']); print('value', d['value'])"
