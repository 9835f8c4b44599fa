"""L2 eviction-policy sweep of the chained decode loop (config 2)."""
import itertools, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sat_b200
B, L, D, H, V, T = 64, 196, 512, 1024, 10000, 20
cfg = sat_b200.Config(batch_size=B, beam_size=1, num_ctx=L, dim_ctx=D, num_lstm_units=H, vocabulary_size=V, max_caption_length=T)
m = sat_b200.CaptionGenerator(cfg)
g = torch.Generator().manual_seed(1)
m.set_weights({n: torch.rand(*s, generator=g) * 0.16 - 0.08 for n, s in sat_b200.weight_shapes(cfg).items()})
pool = [torch.relu(torch.randn(B, L, D, generator=g)).cuda() for _ in range(6)]
def timeit(n=30):
    for i in range(14):
        m.loop_device(pool[i % 6], T)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(m.stream):
        a.record(m.stream)
        for i in range(n):
            m.loop_device(pool[i % 6], T)
        b.record(m.stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n
m.set_option("graphs", 0)   # options are baked into captured graphs: measure eagerly
for stg in (0, 3, 2, 0):
    m.set_option("stages", stg)
    ms = min(timeit() for _ in range(2))
    print("stages=%d : %.3f ms/loop  %.1f us/step" % (stg, ms, ms * 1e3 / T), flush=True)
