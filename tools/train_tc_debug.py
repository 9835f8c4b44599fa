import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_train import setup, TC_DIMS
ocfg, w, m, ctx, sent, masks = setup(B=16, seed=11, dims=TC_DIMS)
res = {}
for tc in (0, 1):
    m.set_option("train_tc", tc)
    m.train_forward_backward(ctx, sent, masks, seed=0)
    res[tc] = {k: v.detach().cpu().numpy().copy() for k, v in m.train_state_dict("grads").items()}
for k in res[0]:
    a, b = res[0][k], res[1][k]
    print("%-28s max|g| %.3e  rel diff %.3e" % (k, np.abs(a).max(), np.abs(a - b).max() / max(np.abs(a).max(), 1e-30)))
