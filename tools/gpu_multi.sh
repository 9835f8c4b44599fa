#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $NG --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_multi.log 2>&1
echo "multi exit $?" >> gpurun_out/bench_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus $NG --impl reference --steps 2 --warmup 1 > gpurun_out/bench_multi_ref.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29535 \
    bench.py --gpus $NG --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train_multi.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -k "train or beam" > gpurun_out/pytest_gpu2.log 2>&1
cat gpurun_out/gpus.txt; tail -n 3 gpurun_out/bench_multi.log | cut -c1-600; tail -n 2 gpurun_out/bench_multi_ref.log | cut -c1-300; tail -n 2 gpurun_out/bench_train_multi.log | cut -c1-500; tail -n 2 gpurun_out/pytest_gpu2.log
