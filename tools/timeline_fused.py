"""GPU timeline of one greedy decode loop (config 2) with the chained dense launch (sat_chain.cu): per launch and
per phase, first CTA start / first CTA through its dependency / last accumulator ready / last arrival (device
globaltimer, eager launches).  python tools/timeline_fused.py [chain=1|0]"""
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
for chain in ([int(a) for a in sys.argv[1:]] or [1, 0]):
    m.set_option("chain", chain)
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
    f = lambda i, k: (int(host[4 * i + k]) - t0) / 1e3 if 0 < int(host[4 * i + k]) < 2 ** 63 else float("nan")
    rows = [(names[i], f(i, 0), f(i, 2), f(i, 3), f(i, 1)) for i in range(n)]
    key = "chain" if chain else "lstm"
    idx = [i for i, r in enumerate(rows) if r[0].startswith(key)]
    print("=== chain=%d: %d timeline entries; steps 6..8 (us since loop start; eager launches)" % (chain, n))
    for nm, a, go, md, b in rows[idx[6]:idx[9]]:
        print("  %-16s start %9.2f  go %9.2f  acc-ready %9.2f  end %9.2f   | wait %5.2f main %6.2f tail %5.2f"
              % (nm, a, go, md, b, go - a, md - go, b - md))
    print("  step period: %.2f us" % ((rows[idx[16]][1] - rows[idx[6]][1]) / 10))
    print("  whole loop: first start -> last end %.2f us" % (np.nanmax([r[4] for r in rows]) - np.nanmin([r[1] for r in rows])))
