#!/bin/bash
# N-GPU runs (N = first argument): the training step with its single collective, and the default line with its train sub-record
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --workload 4 > gpurun_out/bench_train_multi$N.log 2>&1
grep '^{' gpurun_out/bench_train_multi$N.log | tail -1 | cut -c1-700
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_multi$N.log 2>&1
grep '^{' gpurun_out/bench_multi$N.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}); print('train:', {k:v for k,v in (d.get('train') or {}).items() if k!='last_losses'})"
tail -3 gpurun_out/bench_multi$N.log | cut -c1-300
