#!/bin/bash
# First GPU call of the next round: A/B timings of the opt-in training variants and full ncu
# captures of the two [B*L]-row products of attend/fc_1a inside the training step.
mkdir -p gpurun_out
for v in 0 1; do
  echo "== SAT_TRAIN_ATTBWD_WAVE=$v"
  SAT_TRAIN_ATTBWD_WAVE=$v timeout 120 python bench.py --workload 4 --steps 10 --warmup 3 --no-cpu 2>&1 | grep "^{" | cut -c90-250
done
echo "== SAT_TRAIN_PDL=0"
SAT_TRAIN_PDL=0 timeout 120 python bench.py --workload 4 --steps 10 --warmup 3 --no-cpu 2>&1 | grep "^{" | cut -c90-250
# grid 392 = T1 = tanh(drop(ctx) W1a + b) (forward); the 8-CTA-cluster launch with K = B*L = its weight gradient
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lin_umma -s 40 -c 12 \
    -o gpurun_out/prof_train_lin -f python bench.py --steps 1 --warmup 1 --no-cpu --workload 4 > gpurun_out/ncu_train_lin.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:att_bwd_fused -c 1 \
    -o gpurun_out/prof_attbwd -f python bench.py --steps 1 --warmup 1 --no-cpu --workload 4 > gpurun_out/ncu_attbwd.log 2>&1
tail -n 1 gpurun_out/ncu_train_lin.log | cut -c1-200
