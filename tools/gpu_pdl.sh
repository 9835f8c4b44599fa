#!/bin/bash
# scratch script for one-off GPU runs (gpurun -- 'bash tools/gpu_pdl.sh'); the standard rounds are gpu_round.sh / gpu_multi.sh
mkdir -p gpurun_out
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
m.set_option("xbatch", 1)
for sp in (0, 4, 2, 8, 0):
    m.set_option("dec1_splits", sp)
    for i in range(20): m.loop_device(pool[i % 6], T)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(m.stream):
        a.record(m.stream)
        for i in range(40): m.loop_device(pool[i % 6], T)
        b.record(m.stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 40
    print("dec1_splits=%d: %.3f ms/loop %.0f tok/s" % (sp, ms, B * T / ms * 1e3), flush=True)
PY
