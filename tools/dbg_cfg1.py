import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _util import make_pair, rel_err
from oracle import ref_step as R
B, T = 4, 6
ocfg, w, m = make_pair(B)
ctx = R.synth_contexts(ocfg, B)
rng = np.random.RandomState(8)
forced = rng.randint(1, ocfg.vocabulary_size, (B, T)).astype(np.int32)
_, steps = R.decode_loop(ocfg, w, ctx, T, forced, np.float32)
for chain in (0, 1):
    m.set_option("chain", chain)
    for graphs in (0, 1):
        m.set_option("graphs", graphs)
        for rep in range(3 if graphs else 1):
            toks, logits = m.decode_loop(ctx, T, forced, want_logits=True)
        print("chain=%d graphs=%d" % (chain, graphs), " ".join("%.1e" % rel_err(logits[t], steps[t]["logits"]) for t in range(T)))
