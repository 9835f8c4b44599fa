"""Per-K-block cadence of the LSTM tile inside the chained launch: default layout, 128B-swizzle layout, and (timing
experiment, wrong numbers) one MMA per K step instead of the three of the bf16x3 split."""
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
for layout in (0, 1):
    m = sat_b200.CaptionGenerator(cfg)
    m.set_option("umma_layout", layout)
    m.set_weights(W)
    m.set_option("graphs", 0)
    for i in range(3):
        m.loop_device(ctx, T)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m.set_option("graphs", 1)
    for i in range(3):
        m.loop_device(ctx, T)
    with torch.cuda.stream(m.stream):
        a.record(m.stream)
        for i in range(20):
            m.loop_device(ctx, T)
        b.record(m.stream)
    torch.cuda.synchronize()
    print("layout %d: %.1f us per step (graph replay, one context batch)" % (layout, a.elapsed_time(b) / 20 / T * 1e3))
    m.set_option("graphs", 0)
    for mode in (5, 6):
        m.set_option("trace", mode)
        m.set_option("trace_at", 9)
        m.loop_device(ctx, T)
        torch.cuda.synchronize()
        host = np.zeros(1024 * 16, np.int64)
        cr.cudaMemcpy(host.ctypes.data, m.info("trace_ptr"), host.nbytes, cr.cudaMemcpyKind.cudaMemcpyDeviceToHost)
        tr = host.reshape(1024, 16)
        tr = tr[tr[:, 0] > 0]
        t0 = tr[:, 0].min()
        print("== layout %d, %s: phase 0 per K block (us)" % (layout, "three MMAs per K step" if mode == 5 else "ONE MMA per K step (timing experiment)"))
        for i in range(8):
            x, y = (tr[:, 8 + i] - t0) / 1e3, (tr[:, i] - t0) / 1e3
            print("  K block %d: weight copy issued mean %7.2f   operands landed mean %7.2f (min %7.2f max %7.2f)" % (i, x.mean(), y.mean(), y.min(), y.max()))
        m.set_option("trace", 0)
    m.close()
