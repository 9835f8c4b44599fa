#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python tools/timeline.py > gpurun_out/timeline.log 2>&1
grep -A9 "=== overlap=2" gpurun_out/timeline.log | head -10; grep "period" gpurun_out/timeline.log | head -1
for g in 1 0; do
timeout 600 python - <<PY
import sys; sys.path.insert(0, '.')
import torch, sat_b200
B, L, D, H, V, T = 64, 196, 512, 1024, 10000, 20
cfg = sat_b200.Config(batch_size=B, beam_size=1, num_ctx=L, dim_ctx=D, num_lstm_units=H, vocabulary_size=V, max_caption_length=T)
m = sat_b200.CaptionGenerator(cfg)
g = torch.Generator().manual_seed(1)
m.set_weights({n: torch.rand(*s, generator=g) * 0.16 - 0.08 for n, s in sat_b200.weight_shapes(cfg).items()})
pool = [torch.relu(torch.randn(B, L, D, generator=g)).cuda() for _ in range(6)]
torch.cuda.synchronize()
m.set_option("gate", $g); m.set_option("xbatch", 1)
for i in range(20): m.loop_device(pool[i % 6], T)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(m.stream):
    a.record(m.stream)
    for i in range(40): m.loop_device(pool[i % 6], T)
    b.record(m.stream)
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 40
print("gate=$g: %.3f ms/loop %.0f tok/s" % (ms, B * T / ms * 1e3))
PY
done
