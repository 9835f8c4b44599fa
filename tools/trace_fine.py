"""Fine stamps of K blocks 0..3 of the LSTM tile in the chained launch: activation copy issued / weights seen / activations
seen / MMAs committed."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sat_b200
import cuda.bindings.runtime as cr
B, L, D, H, V, T = 64, 196, 512, 1024, 10000, 20
cfg = sat_b200.Config(batch_size=B, beam_size=1, num_ctx=L, dim_ctx=D, num_lstm_units=H, vocabulary_size=V, max_caption_length=T)
g = torch.Generator().manual_seed(1)
W = {n: torch.rand(*s, generator=g) * 0.16 - 0.08 for n, s in sat_b200.weight_shapes(cfg).items()}
ctx = torch.relu(torch.randn(B, L, D, generator=g)).cuda()
m = sat_b200.CaptionGenerator(cfg)
m.set_weights(W)
m.set_option("graphs", 0)
for i in range(3):
    m.loop_device(ctx, T)
torch.cuda.synchronize()
m.set_option("trace", 7)
m.set_option("trace_at", 9)
m.loop_device(ctx, T)
torch.cuda.synchronize()
host = np.zeros(1024 * 16, np.int64)
cr.cudaMemcpy(host.ctypes.data, m.info("trace_ptr"), host.nbytes, cr.cudaMemcpyKind.cudaMemcpyDeviceToHost)
tr = host.reshape(1024, 16)
tr = tr[tr[:, 4] > 0]
t0 = tr[:, 0].min()
names = ["X copy issued", "X seen by MMA lane", "MMAs committed", "W seen by MMA lane"]
for blk in range(4):
    print("K block %d:" % blk, "  ".join("%s %6.2f (min %6.2f)" % (names[j], ((tr[:, 4 * j + blk] - t0) / 1e3).mean(), ((tr[:, 4 * j + blk] - t0) / 1e3).min())
                                          for j in (0, 3, 1, 2)))
for c in (0, 1, 40, 127):
    if c < len(tr):
        print("CTA row %d:" % c, [round((int(v) - int(t0)) / 1e3, 2) for v in tr[c]])
