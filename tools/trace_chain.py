"""In-kernel stamps of ONE chained dense launch (sat_chain.cu) inside the eager greedy loop at config 2: where a step's
time goes CTA by CTA.  python tools/trace_chain.py [launch numbers ...]"""
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
labels = {0: "CTA start", 1: "TMA lane: dependency wait returned (phase 0 open)", 2: "ph0 MMA: first operands landed",
          3: "ph0 MMA: all issued", 4: "ph0 epilogue: accumulator ready", 5: "CTA end",
          6: "ph0 epilogue: split-K rendezvous passed", 7: "ph0 epilogue: rows done", 8: "ph0 arrived",
          9: "ph1 MMA: first operands landed", 10: "ph1 epilogue: accumulator ready", 15: "ph1 arrived",
          11: "ph2 MMA: first operands landed", 12: "ph2 epilogue: accumulator ready", 13: "ph2 arrived",
          14: "ph2 last arriver: words recorded"}
order = [0, 1, 2, 3, 4, 6, 7, 8, 9, 10, 15, 11, 12, 13, 14, 5]
m.set_option("graphs", 0)
for i in range(3):
    m.loop_device(ctx, T)
torch.cuda.synchronize()
for k in ([int(a) for a in sys.argv[1:]] or [7, 12]):
    m.set_option("trace", 4)
    m.set_option("trace_at", k)
    m.loop_device(ctx, T)
    torch.cuda.synchronize()
    host = np.zeros(1024 * 16, np.int64)
    cr.cudaMemcpy(host.ctypes.data, m.info("trace_ptr"), host.nbytes, cr.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    tr = host.reshape(1024, 16)
    tr = tr[tr[:, 0] > 0]
    t0 = tr[:, 1][tr[:, 1] > 0].min()
    print("== chained launch #%d of the loop: %d CTAs (us relative to the first CTA through its dependency wait)" % (k, len(tr)))
    for i in order:
        col = tr[:, i]; ok = col > 0
        if ok.any():
            v = (col[ok] - t0) / 1e3
            print("  %-52s mean %7.2f  min %7.2f  max %7.2f   (n=%d)" % (labels[i], v.mean(), v.min(), v.max(), ok.sum()))
    m.set_option("trace", 0)
# per-K-block view of phase 0 (LSTM tile): when each weight copy was issued and when the block's operands had landed
for k in [9]:
    m.set_option("trace", 5)
    m.set_option("trace_at", k)
    m.loop_device(ctx, T)
    torch.cuda.synchronize()
    host = np.zeros(1024 * 16, np.int64)
    cr.cudaMemcpy(host.ctypes.data, m.info("trace_ptr"), host.nbytes, cr.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    tr = host.reshape(1024, 16)
    tr = tr[tr[:, 0] > 0]
    t0 = tr[:, 0].min()
    print("== chained launch #%d, phase 0 per K block (us relative to the first block-0 operands landed)" % k)
    for i in range(8):
        a, b = (tr[:, 8 + i] - t0) / 1e3, (tr[:, i] - t0) / 1e3
        print("  K block %d: weight copy issued mean %7.2f (min %7.2f max %7.2f)   operands landed mean %7.2f (min %7.2f max %7.2f)"
              % (i, a.mean(), a.min(), a.max(), b.mean(), b.min(), b.max()))
    m.set_option("trace", 0)
