"""GPU timeline of the beam-search loop (config 5): dense / attention launches of two steps."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sat_b200
import cuda.bindings.runtime as cr
B, L, D, H, V, T, beam = 128, 196, 512, 1024, 10000, 30, 3
cfg = sat_b200.Config(batch_size=B, beam_size=beam, num_ctx=L, dim_ctx=D, num_lstm_units=H, vocabulary_size=V, max_caption_length=T)
m = sat_b200.CaptionGenerator(cfg)
g = torch.Generator().manual_seed(1)
m.set_weights({n: torch.rand(*s, generator=g) * 0.16 - 0.08 for n, s in sat_b200.weight_shapes(cfg).items()})
ctx = torch.relu(torch.randn(B, L, D, generator=g)).cuda()
m.set_option("graphs", 0)
for i in range(3):
    m.beam_device(ctx, beam, T, 2)
torch.cuda.synchronize()
m.set_option("trace", 3)
m.beam_device(ctx, beam, T, 2)
torch.cuda.synchronize()
n = m.info("tl_count")
host = np.zeros(1024 * 16, np.uint64)
cr.cudaMemcpy(host.ctypes.data, m.info("trace_ptr"), host.nbytes, cr.cudaMemcpyKind.cudaMemcpyDeviceToHost)
names = []
for i in range(n):
    m.info("tl_tag_%d" % i); names.append(m.lib.sat_last_error().decode())
m.set_option("trace", 0)
t0 = int(host[0])
f = lambda i, k: (int(host[4 * i + k]) - t0) / 1e3
rows = [(names[i], f(i, 0), f(i, 2), f(i, 3), f(i, 1)) for i in range(n)]
lstm_idx = [i for i, r in enumerate(rows) if r[0].startswith("lstm")]
for nm, a, go, md, b in rows[lstm_idx[10]:lstm_idx[12]]:
    print("  %-18s start %9.2f  go %9.2f  main-done %9.2f  end %9.2f   | prologue %5.2f main %6.2f tail %5.2f" % (nm, a, go, md, b, go - a, md - go, b - md))
print("  step period: %.2f us" % ((rows[lstm_idx[20]][1] - rows[lstm_idx[10]][1]) / 10))
m.set_option("profile", 1)
m.beam_device(ctx, beam, T, 2)
torch.cuda.synchronize()
for t in ("att_state", "att", "lstm", "dec1", "dec2", "rows", "beam", "proj", "init"):
    n = m.info("prof_n_" + t)
    if n: print("  profile (events around each launch, eager): %-10s n=%3d mean %.2f us" % (t, n, m.info("prof_ns_" + t) / n / 1e3))
m.set_option("profile", 0)
