#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 900 python bench.py --workload 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_train1.log 2>&1
if [ "$NG" -gt 1 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29535 \
    bench.py --gpus $NG --workload 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_train_multi.log 2>&1
fi
tail -n 2 gpurun_out/bench_train1.log | cut -c1-900; [ -f gpurun_out/bench_train_multi.log ] && tail -n 3 gpurun_out/bench_train_multi.log | cut -c1-900
