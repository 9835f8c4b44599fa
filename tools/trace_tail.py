"""Epilogue internals of the chained launch: the split-K tail of phase 0 and the last arriver's tail of phase 2."""
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
m.set_option("trace", 8)
m.set_option("trace_at", 9)
m.loop_device(ctx, T)
torch.cuda.synchronize()
host = np.zeros(1024 * 16, np.int64)
cr.cudaMemcpy(host.ctypes.data, m.info("trace_ptr"), host.nbytes, cr.cudaMemcpyKind.cudaMemcpyDeviceToHost)
tr = host.reshape(1024, 16)
tr = tr[tr[:, 5] > 0]
t0 = tr[:, 6][tr[:, 6] > 0].min()
lab = {6: "ph0 accumulator ready", 7: "ph0 tile parked (smem + scratch stores issued)", 8: "ph0 fence + barrier done",
       9: "ph0 split-K rendezvous passed", 10: "ph0 first row: partials in registers", 11: "ph0 rows done", 12: "ph0 arrived",
       0: "LAST ARRIVER: arrival returned", 1: "LAST ARRIVER: words recorded", 2: "LAST ARRIVER: row copies issued",
       3: "LAST ARRIVER: rows landed", 4: "LAST ARRIVER: conversion done", 5: "CTA end"}
for i in (6, 7, 8, 9, 10, 11, 12, 0, 1, 2, 3, 4, 5):
    col = tr[:, i]; ok = col > 0
    if ok.any():
        v = (col[ok] - t0) / 1e3
        print("  %-48s mean %7.2f  min %7.2f  max %7.2f   (n=%d)" % (lab[i], v.mean(), v.min(), v.max(), ok.sum()))
