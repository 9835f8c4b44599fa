"""GPU bring-up: run each kernel family in its own subprocess (a trap or hang in one does not hide the
others) and log max-norm errors against numpy.  Usage under gpurun:
    python tools/bringup.py > gpurun_out/bringup.log 2>&1
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DENSE = r'''
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import sat_b200
cfg = sat_b200.Config(batch_size=4, beam_size=1, num_ctx=49, dim_ctx=64, dim_embedding=32, num_lstm_units=64,
                      dim_initalize_layer=32, dim_attend_layer=32, dim_decode_layer=64, vocabulary_size=300)
m = sat_b200.CaptionGenerator(cfg)
m.set_option("umma_layout", %(layout)d); m.set_option("gemm", %(gemm)d)
for rows, K, n, act, sp in [(4,64,128,0,1),(4,128,128,0,2),(16,64,128,0,1),(64,2048,4096,0,0),(64,1024,10000,0,0),(3,72,50,1,1),(200,512,512,1,1),(384,2048,1024,1,0)]:
    rng = np.random.RandomState(0)
    x = rng.uniform(-1,1,(rows,K)).astype(np.float32); w = rng.uniform(-.08,.08,(K,n)).astype(np.float32); b = rng.uniform(-.08,.08,(n,)).astype(np.float32)
    xd, wd, bd = (torch.from_numpy(a).cuda() for a in (x,w,b)); y = torch.full((rows,n), float("nan"), device="cuda")
    torch.cuda.synchronize(); p = lambda t: C.c_void_p(t.data_ptr())
    rc = m.lib.sat_dense_fwd(m._h, p(xd), p(wd), p(bd), p(y), rows, K, n, act, sp, m._st())
    torch.cuda.synchronize()
    ref = x.astype(np.float64) @ w.astype(np.float64) + b
    if act: ref = np.tanh(ref)
    got = y.cpu().numpy()
    err = np.abs(got-ref).max()/np.abs(ref).max()
    print("dense layout=%(layout)d gemm=%(gemm)d rows=%%d K=%%d n=%%d act=%%d splits=%%d rc=%%d relerr=%%.3e nan=%%d" %% (rows,K,n,act,sp,rc,err,int(np.isnan(got).sum())), flush=True)
'''

STEP = r'''
import sys, numpy as np, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import sat_b200
from oracle import ref_step as R
from _util import make_pair, rel_err, SMALL
layers = %(layers)d
dims = dict(SMALL) if %(small)d else {}
ocfg, w, m = make_pair(4, num_attend_layers=layers, num_decode_layers=layers, num_initalize_layers=layers, **dims)
m.set_option("gemm", %(gemm)d)
H, V = ocfg.num_lstm_units, ocfg.vocabulary_size
ctx = R.synth_contexts(ocfg, 4); rng = np.random.RandomState(0)
lw = rng.randint(0, V, 4).astype(np.int32); c = rng.uniform(-.5,.5,(4,H)).astype(np.float32); h = rng.uniform(-.5,.5,(4,H)).astype(np.float32)
ref = R.decode_step(ocfg, w, ctx, lw, c, h, np.float64)
c0, h0 = m.initialize(ctx); rc0, rh0 = R.initialize(ocfg, w, ctx, np.float64)
print("init  layers=%%d small=%(small)d gemm=%(gemm)d c0=%%.3e h0=%%.3e" %% (layers, rel_err(c0, rc0), rel_err(h0, rh0)), flush=True)
got = m.decode_step(ctx, lw, c, h, extras=True)
print("step  layers=%%d small=%(small)d gemm=%(gemm)d " %% layers + " ".join("%%s=%%.3e" %% (k, rel_err(got[k], ref[k])) for k in ("alpha","memory","output","logits","probs")), flush=True)
toks, lg = m.decode_loop(ctx, 4, None, want_logits=True)
rt, steps = R.decode_loop(ocfg, w, ctx, 4, None, np.float64)
print("loop  logits3=%%.3e tokens_equal=%%s" %% (rel_err(lg[3], steps[3]["logits"]), bool((toks==rt).all())), flush=True)
toks2, lg2 = m.decode_loop(ctx, 4, None, want_logits=True)
toks3, lg3 = m.decode_loop(ctx, 4, None, want_logits=True)
print("graph replay identical:", bool(np.array_equal(lg, lg2) and np.array_equal(lg2, lg3)), flush=True)
'''


def run(tag, code, timeout=240):
    print("=== %s" % tag, flush=True)
    try:
        r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                           timeout=timeout)
        out = r.stdout.strip().splitlines()
        print("\n".join(out[-25:]))
        print("--- exit %d" % r.returncode, flush=True)
        return r.returncode
    except subprocess.TimeoutExpired as e:
        print((e.stdout or b"").decode(errors="replace")[-2000:] if isinstance(e.stdout, bytes) else (e.stdout or ""))
        print("--- TIMEOUT", flush=True)
        return -1


if __name__ == "__main__":
    for layout in (0, 1):
        for gemm in (0, 1):
            run("dense layout=%d gemm=%d" % (layout, gemm), DENSE % dict(root=ROOT, layout=layout, gemm=gemm))
    for gemm in (0, 1):
        for small in (1, 0):
            for layers in (2, 1):
                run("step layers=%d small=%d gemm=%d" % (layers, small, gemm),
                    STEP % dict(root=ROOT, layers=layers, small=small, gemm=gemm))
