"""Per-CTA start/end of the last attention launch inside the decode loop (trace mode 2)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sat_b200
import cuda.bindings.runtime as cr
B, L, D, H, V, T = 64, 196, 512, 1024, 10000, 20
cfg = sat_b200.Config(batch_size=B, beam_size=1, num_ctx=L, dim_ctx=D, num_lstm_units=H, vocabulary_size=V, max_caption_length=T)
m = sat_b200.CaptionGenerator(cfg)
g = torch.Generator().manual_seed(1)
m.set_weights({n: torch.rand(*s, generator=g) * 0.16 - 0.08 for n, s in sat_b200.weight_shapes(cfg).items()})
ctx = torch.relu(torch.randn(B, L, D, generator=g)).cuda()
for overlap in (2, 1):
    m.set_option("overlap", overlap); m.set_option("pdl", 1); m.set_option("graphs", 0)
    for i in range(3):
        m.loop_device(ctx, T)
    torch.cuda.synchronize()
    m.set_option("trace", 2)
    m.loop_device(ctx, T)
    torch.cuda.synchronize()
    host = np.zeros(1024 * 16, np.int64)
    cr.cudaMemcpy(host.ctypes.data, m.info("trace_ptr"), host.nbytes, cr.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    m.set_option("trace", 0)
    tr = host.reshape(1024, 16)
    tr = tr[tr[:, 0] > 0]
    t0 = tr[:, 0].min()
    st = np.sort((tr[:, 0] - t0) / 1e3); en = np.sort((tr[:, 7] - t0) / 1e3)
    dur = (tr[:, 7] - tr[:, 0]) / 1e3
    print("overlap=%d: %d CTAs; start pct [0,25,50,75,100] = %s; end = %s; per-CTA duration min/mean/max = %.1f %.1f %.1f"
          % (overlap, len(tr), np.percentile(st, [0, 25, 50, 75, 100]).round(1), np.percentile(en, [0, 25, 50, 75, 100]).round(1),
             dur.min(), dur.mean(), dur.max()))
    for i in range(2, 7):
        col = tr[:, i]; ok = col > 0
        if ok.any(): print("   stamp %d mean %.2f max %.2f" % (i, ((col[ok] - t0) / 1e3).mean(), ((col[ok] - t0) / 1e3).max()))
