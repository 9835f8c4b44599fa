import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _util import make_pair, rel_err
from oracle import ref_step as R
chain = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B, T = 64, 4
ocfg, w, m = make_pair(B, num_lstm_units=1024, vocabulary_size=10000)
m.set_option("chain", chain)
m.set_option("graphs", 0)
ctx = R.synth_contexts(ocfg, B)
rng = np.random.RandomState(2)
forced = rng.randint(1, 10000, (B, T)).astype(np.int32)
_, steps = R.decode_loop(ocfg, w, ctx, T, forced, np.float32)
c0, h0 = m.initialize(ctx)
c0r, h0r = R.initialize(ocfg, w, ctx)
print("initialize: c0 err %.2e h0 err %.2e" % (rel_err(c0, c0r), rel_err(h0, h0r)))
toks, logits = m.decode_loop(ctx, T, forced, want_logits=True)
for t in range(T):
    print("chain=%d step %d logits err %.3e nan=%d" % (chain, t, rel_err(np.nan_to_num(logits[t]), steps[t]["logits"]), int(np.isnan(logits[t]).sum())))
