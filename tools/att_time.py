"""How long does the attention kernel take? CUDA events (as bench.py does) vs device globaltimer stamps."""
import os, sys, ctypes as C
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
hstate = (torch.rand(B, H, generator=g) - 0.5).cuda()
alpha = torch.empty(B, L, device="cuda"); z = torch.empty(B, D, device="cuda")
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
m.prepare(ctx, want_state=False)
st = m.stream
def run(flushit):
    with torch.cuda.stream(st):
        if flushit: flush.zero_()
        assert m.lib.sat_attention_fwd(m._h, p(ctx), p(hstate), p(alpha), p(z), B, 1, m._st()) == 0
for sms in (0, 64):
    for pdl in (1, 0):
        m.set_option("att_sms", sms); m.set_option("pdl", pdl)
        for flushit in (True, False):
            for i in range(3): run(flushit)
            torch.cuda.synchronize()
            m.set_option("profile", 1)
            for i in range(20): run(flushit)
            torch.cuda.synchronize()
            ev = m.info("prof_ns_att") / max(1, m.info("prof_n_att")) / 1e3
            m.set_option("profile", 0)
            # device stamps: trace mode 3 records {min start, max end, go, main done} of every launch
            m.set_option("trace", 3)
            for i in range(6): run(flushit)
            torch.cuda.synchronize()
            n = m.info("tl_count")
            host = np.zeros(1024 * 16, np.uint64)
            cr.cudaMemcpy(host.ctypes.data, m.info("trace_ptr"), host.nbytes, cr.cudaMemcpyKind.cudaMemcpyDeviceToHost)
            names = []
            for i in range(n):
                m.info("tl_tag_%d" % i); names.append(m.lib.sat_last_error().decode())
            m.set_option("trace", 0)
            d = [(int(host[4*i+1]) - int(host[4*i])) / 1e3 for i in range(n) if names[i].startswith("attention")]
            dg = [(int(host[4*i+1]) - int(host[4*i+2])) / 1e3 for i in range(n) if names[i].startswith("attention")]
            print("att_sms=%3d pdl=%d flush=%d: events %.2f us | device first-CTA-start -> last-CTA-end %.2f us, after-wait -> end %.2f us (median of %d)"
                  % (sms, pdl, flushit, ev, np.median(d), np.median(dg), len(d)))
