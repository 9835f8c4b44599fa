"""In-kernel stamps of the vocabulary layer (decode fc_2 with the fused argmax) inside the eager decode loop.
    python tools/trace_loop.py"""
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
labels = ["start", "producers start", "row loop: sums + bias ready (last row)", "row loop: activation done (last row)", "MMA: first W stage",
          "MMA: first X stage", "MMA: all issued", "epilogue: accumulator ready", "partials written",
          "rendezvous passed", "end", "am: candidates stored", "am: grid barrier passed", "am: rows of this CTA finished", "-", "row loop begins"]
want = set(int(a) for a in sys.argv[1:]) or {79}
for overlap, warm in ((2, 1),):
    m.set_option("overlap", overlap)
    m.set_option("warm", warm)
    m.set_option("graphs", 0)
    for i in range(3):
        m.loop_device(ctx, T)
    torch.cuda.synchronize()
    seen = set()
    for k in list(range(0, 3)) + list(range(10, 40)):
        m.set_option("trace", 1)
        m.set_option("trace_at", k)
        m.loop_device(ctx, T)
        torch.cuda.synchronize()
        host = np.zeros(1024 * 16, np.int64)
        cr.cudaMemcpy(host.ctypes.data, m.info("trace_ptr"), host.nbytes, cr.cudaMemcpyKind.cudaMemcpyDeviceToHost)
        tr = host.reshape(1024, 16)
        tr = tr[tr[:, 0] > 0]
        n = len(tr)
        if n not in want or n in seen or (3 <= k < 14):
            continue
        seen.add(n)
        t0 = tr[:, 0].min()
        print("== overlap=%d dense launch #%d: %d CTAs (us after first CTA start)" % (overlap, k, n))
        for i, lab in enumerate(labels):
            col = tr[:, i]; ok = col > 0
            if ok.any():
                v = (col[ok] - t0) / 1e3
                print("  %-34s mean %7.2f  min %7.2f  max %7.2f   (n=%d)" % (lab, v.mean(), v.min(), v.max(), ok.sum()))
    m.set_option("trace", 0)
