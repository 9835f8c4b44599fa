"""GPU timeline of one decode loop (config 2): per kernel launch, first CTA start and last CTA end.
    python tools/timeline.py [overlap]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sat_b200
B, L, D, H, V, T = 64, 196, 512, 1024, 10000, 20
cfg = sat_b200.Config(batch_size=B, beam_size=1, num_ctx=L, dim_ctx=D, num_lstm_units=H, vocabulary_size=V, max_caption_length=T)
m = sat_b200.CaptionGenerator(cfg)
g = torch.Generator().manual_seed(1)
m.set_weights({n: torch.rand(*s, generator=g) * 0.16 - 0.08 for n, s in sat_b200.weight_shapes(cfg).items()})
ctx = torch.relu(torch.randn(B, L, D, generator=g)).cuda()
import cuda.bindings.runtime as cr
for overlap, pdl in ((2, 1), (1, 1)):
    m.set_option("overlap", overlap)
    m.set_option("pdl", pdl)
    m.set_option("graphs", 0)
    for i in range(3):
        m.loop_device(ctx, T)
    torch.cuda.synchronize()
    m.set_option("trace", 3)
    m.loop_device(ctx, T)
    torch.cuda.synchronize()
    n = m.info("tl_count")
    host = np.zeros(1024 * 16, np.uint64)
    cr.cudaMemcpy(host.ctypes.data, m.info("trace_ptr"), host.nbytes, cr.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    names = []
    for i in range(n):
        m.info("tl_tag_%d" % i)
        names.append(m.lib.sat_last_error().decode())
    m.set_option("trace", 0)
    t0 = int(host[0])
    print("=== overlap=%d pdl=%d: %d launches; 3 steps (us since loop start; eager launches)" % (overlap, pdl, n))
    f = lambda i, k: (int(host[4 * i + k]) - t0) / 1e3
    rows = [(names[i], f(i, 0), f(i, 2), f(i, 3), f(i, 1)) for i in range(n)]
    if "--head" in sys.argv:
        for nm, a, go, md, b in rows[:14]:
            print("  %-18s start %9.2f  go %9.2f  main-done %9.2f  end %9.2f   | prologue %5.2f main %6.2f tail %5.2f"
                  % (nm, a, go, md, b, go - a, md - go, b - md))
        print("  ...")
    lstm_idx = [i for i, r in enumerate(rows) if r[0].startswith("lstm")]
    lo, hi = lstm_idx[5], lstm_idx[8]
    for nm, a, go, md, b in rows[lo:hi]:
        print("  %-18s start %9.2f  go %9.2f  main-done %9.2f  end %9.2f   | prologue %5.2f main %6.2f tail %5.2f"
              % (nm, a, go, md, b, go - a, md - go, b - md))
    print("  step period: %.2f us" % ((rows[lstm_idx[15]][1] - rows[lstm_idx[5]][1]) / 10))
    print("  whole loop: first start -> last end %.2f us" % (max(r[4] for r in rows) - min(r[1] for r in rows)))
