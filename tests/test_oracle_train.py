"""The autograd training oracle (oracle/train_ref.py) against the numpy restatement (oracle/ref_step.py)
and against finite differences; the dropout RNG is deterministic and has the right rate."""
import numpy as np

from oracle import ref_step as R
from oracle import train_ref as TR


def small_cfg():
    return R.OracleConfig(num_ctx=9, dim_ctx=16, dim_embedding=8, num_lstm_units=16, dim_initalize_layer=8,
                          dim_attend_layer=8, dim_decode_layer=16, vocabulary_size=30, batch_size=3,
                          max_caption_length=4)


def data(cfg, seed=0):
    rng = np.random.RandomState(seed)
    w = R.init_weights(cfg, seed)
    ctx = R.synth_contexts(cfg, 3, seed)
    sent = rng.randint(1, cfg.vocabulary_size, (3, 4)).astype(np.int32)
    masks = (np.arange(4)[None, :] < np.array([4, 2, 3])[:, None]).astype(np.float32)
    return w, ctx, sent, masks


def test_rng_is_deterministic_and_has_the_keep_rate():
    a = TR.dropout_mask(7, 3, (1000, 100), 0.7)
    b = TR.dropout_mask(7, 3, (1000, 100), 0.7)
    assert np.array_equal(a, b) and set(np.unique(a)) == {0.0, 1.0}
    assert abs(a.mean() - 0.7) < 5e-3
    assert abs(TR.dropout_mask(7, 4, (1000, 100), 0.5).mean() - 0.5) < 5e-3
    assert not np.array_equal(a, TR.dropout_mask(8, 3, (1000, 100), 0.7))
    # first values pinned (the CUDA generator must reproduce them: tests/test_gpu_train.py)
    u = TR.uniform24(1234, 5, 4)
    assert u.dtype == np.float32 and (u >= 0).all() and (u < 1).all()


def test_torch_forward_equals_numpy_forward_with_and_without_dropout():
    cfg = small_cfg()
    w, ctx, sent, masks = data(cfg)
    for seed in (None, 11):
        dm = [TR.step_masks(cfg, seed, t, 3) for t in range(4)] if seed is not None else None
        im = TR.init_masks(cfg, seed, 3) if seed is not None else None
        ref = R.train_forward(cfg, w, ctx, sent, masks, np.float64, dm, im)
        got, _ = TR.loss_and_grads(cfg, w, ctx, sent, masks, seed)
        for k in ("total_loss", "cross_entropy_loss", "attention_loss", "reg_loss", "accuracy"):
            assert abs(got[k] - ref[k]) < 1e-10 * max(1.0, abs(ref[k])), (seed, k)


def test_gradients_match_finite_differences():
    cfg = small_cfg()
    w, ctx, sent, masks = data(cfg, 1)
    _, g = TR.loss_and_grads(cfg, w, ctx, sent, masks, seed=5)
    rng = np.random.RandomState(0)
    for name in ("lstm/lstm_cell/kernel", "attend/fc_1a/kernel", "attend/fc_2/kernel", "decode/fc_2/bias",
                 "word_embedding/weights", "initialize/fc_b2/kernel"):
        for _ in range(2):
            idx = tuple(rng.randint(0, s) for s in w[name].shape)
            eps = 1e-5
            wp = {k: v.astype(np.float64).copy() for k, v in w.items()}
            wm = {k: v.astype(np.float64).copy() for k, v in w.items()}
            wp[name][idx] += eps
            wm[name][idx] -= eps
            lp, _ = TR.loss_and_grads(cfg, wp, ctx, sent, masks, seed=5)
            lm, _ = TR.loss_and_grads(cfg, wm, ctx, sent, masks, seed=5)
            fd = (lp["total_loss"] - lm["total_loss"]) / (2 * eps)
            assert abs(fd - g[name][idx]) < 1e-6 + 1e-4 * abs(fd), (name, idx, fd, g[name][idx])


def test_data_parallel_shards_sum_to_the_global_gradient():
    """Two shards with global normalisers + the regulariser added once == the single-process gradient
    (CE by the global mask sum, coverage loss by the global batch)."""
    cfg = small_cfg()
    w, ctx, sent, masks = data(cfg, 2)
    full_l, full_g = TR.loss_and_grads(cfg, w, ctx, sent, masks, seed=None)
    msum = float(masks.sum())
    parts = []
    for lo, hi in ((0, 2), (2, 3)):
        l, g = TR.loss_and_grads(cfg, w, ctx[lo:hi], sent[lo:hi], masks[lo:hi], None, msum, 3, reg_in_grad=False)
        parts.append((l, g))
    ce = sum(p[0]["cross_entropy_loss"] for p in parts)
    assert abs(ce - full_l["cross_entropy_loss"]) < 1e-10
    reg_names = set(R.regularized_names(w))
    for k in w:
        tot = parts[0][1][k] + parts[1][1][k]
        if k in reg_names:
            tot = tot + cfg.fc_kernel_regularizer_scale * w[k].astype(np.float64)
        np.testing.assert_allclose(tot, full_g[k], rtol=1e-9, atol=1e-12)


def test_clip_and_adam_first_step_moves_by_lr():
    w = {"a": np.ones((3, 3)), "b": np.full((2,), 2.0)}
    g = {"a": np.full((3, 3), 10.0), "b": np.full((2,), -10.0)}
    m = {k: np.zeros_like(v) for k, v in w.items()}
    v = {k: np.zeros_like(x) for k, x in w.items()}
    nw, nm, nv, norm = TR.clip_and_adam(w, g, m, v, step=1)
    assert abs(norm - np.sqrt(11 * 100.0)) < 1e-9
    # first Adam step: |dw| = lr * |g|/(|g| + eps*sqrt(1-b2)) ~ lr
    assert np.allclose(nw["a"], 1.0 - 1e-4, atol=1e-8) and np.allclose(nw["b"], 2.0 + 1e-4, atol=1e-8)
