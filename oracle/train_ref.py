"""Training-step oracle: forward + backward + clip + Adam of the unrolled training graph
(model.py:250-334 losses, :461-511 optimizer) as a torch-CPU float64 restatement with autograd.

TEST INFRASTRUCTURE ONLY (same rules as oracle/ref_step.py; PARITY UNPINNED for the same reason:
TensorFlow cannot run here).  torch is used for autograd only; the arithmetic mirrors
oracle/ref_step.train_forward line by line, and tests/test_oracle_train.py checks the two against
each other.

Dropout masks are INJECTED: both this oracle and the CUDA path derive them from the counter-based
generator `dropout_mask` below (the reference uses unseeded TF random ops), so a
training step is reproducible bit for bit on both sides.
"""
from __future__ import annotations

import numpy as np

from . import ref_step as R

# ------------------------------------------------------------------ counter-based dropout RNG
# u = 24 high bits of a 32-bit integer hash of idx keyed by K = seed ^ stream*GOLDEN (low word xor-ed in before the
# first multiply, high word added between the two multiplies); mask = floor(keep + u) in float32.  The CUDA side
# (sat_linear.cuh: rng_bits24 / rng_u24) is the same function.
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M32 = np.uint64(0xFFFFFFFF)

# mask streams of time step t: t * 16 + k
ATT_CTX, ATT_OUT, ATT_MID, LSTM_IN, LSTM_STATE, LSTM_OUT, DEC_IN, DEC_MID = range(8)
INIT_BASE = 0xFFFF0   # + 0 init_mean, + 1 init_a, + 2 init_b


def uniform24(seed: int, stream: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        K = np.uint64(seed) ^ (np.uint64(stream) * _GOLD)
        k0, k1 = np.uint32(K & _M32), np.uint32(K >> np.uint64(32))
        idx = np.arange(n, dtype=np.uint64)
        lo, hi = (idx & _M32).astype(np.uint32), (idx >> np.uint64(32)).astype(np.uint32)
        x = lo ^ k0
        x ^= x >> np.uint32(16)
        x *= np.uint32(0x21F0AAAD)
        x ^= x >> np.uint32(15)
        x += k1 ^ (hi * np.uint32(0x9E3779B1))
        x *= np.uint32(0x735A2D97)
        x ^= x >> np.uint32(15)
    return ((x >> np.uint32(8)).astype(np.float32)) * np.float32(2.0 ** -24)


def dropout_mask(seed: int, stream: int, shape, keep: float) -> np.ndarray:
    """0/1 mask = floor(keep + U[0,1)) (graph fixture */dropout/{Floor}), float32 arithmetic."""
    n = int(np.prod(shape))
    u = uniform24(seed, stream, n)
    # (min: with keep == 1 the float32 sum can round up to 2.0 for the largest u)
    return np.minimum(np.floor(np.float32(keep) + u), 1.0).astype(np.float32).reshape(shape)


def step_masks(cfg: R.OracleConfig, seed: int, t: int, B: int):
    """The eight injected masks of time step t (names as in ref_step.train_forward)."""
    L, D, E, H = cfg.num_ctx, cfg.dim_ctx, cfg.dim_embedding, cfg.num_lstm_units
    A, Dd = cfg.dim_attend_layer, cfg.dim_decode_layer
    kf, kl = 1.0 - cfg.fc_drop_rate, 1.0 - cfg.lstm_drop_rate
    s = t * 16
    return dict(att_ctx=dropout_mask(seed, s + ATT_CTX, (B * L, D), kf),
                att_out=dropout_mask(seed, s + ATT_OUT, (B, H), kf),
                att_mid=dropout_mask(seed, s + ATT_MID, (B * L, A), kf),
                lstm_in=dropout_mask(seed, s + LSTM_IN, (B, D + E), kl),
                lstm_state=dropout_mask(seed, s + LSTM_STATE, (B, H), kl),
                lstm_out=dropout_mask(seed, s + LSTM_OUT, (B, H), kl),
                dec_in=dropout_mask(seed, s + DEC_IN, (B, H + D + E), kf),
                dec_mid=dropout_mask(seed, s + DEC_MID, (B, Dd), kf))


def init_masks(cfg: R.OracleConfig, seed: int, B: int):
    kf = 1.0 - cfg.fc_drop_rate
    return dict(init_mean=dropout_mask(seed, INIT_BASE + 0, (B, cfg.dim_ctx), kf),
                init_a=dropout_mask(seed, INIT_BASE + 1, (B, cfg.dim_initalize_layer), kf),
                init_b=dropout_mask(seed, INIT_BASE + 2, (B, cfg.dim_initalize_layer), kf))


# ------------------------------------------------------------------ forward with autograd (2-layer modes)
def forward_torch(cfg: R.OracleConfig, w, contexts, sentences, masks, seed=None,
                  global_mask_sum=None, global_batch=None):
    """Mirror of ref_step.train_forward on torch float64 tensors (1- and 2-layer attend / decode / initialize).

    w: dict name -> torch tensor (requires_grad).  seed=None switches all dropout off.
    global_mask_sum / global_batch: normalisers of the GLOBAL batch when this is one data-parallel shard
    (model.py:316-318 divides by reduce_sum(masks) of the whole batch, :324-326 by batch_size*num_ctx).
    """
    import torch
    f64 = torch.float64
    B, T = sentences.shape
    L = cfg.num_ctx
    kf, kl = 1.0 - cfg.fc_drop_rate, 1.0 - cfg.lstm_drop_rate
    ctx = torch.as_tensor(contexts, dtype=f64)
    mk = torch.as_tensor(masks, dtype=f64)
    sent = torch.as_tensor(np.asarray(sentences), dtype=torch.long)

    def drop(x, m, keep):
        if m is None:
            return x
        return x / keep * torch.as_tensor(m, dtype=f64)

    def dense(x, name, act=None):
        y = x @ w[name + "/kernel"] + w[name + "/bias"]
        return act(y) if act else y

    im = init_masks(cfg, seed, B) if seed is not None else {}
    m = drop(ctx.mean(dim=1), im.get("init_mean"), kf)
    if cfg.num_initalize_layers == 1:                      # model.py:362-371
        c = dense(m, "initialize/fc_a")
        h_state = dense(m, "initialize/fc_b")
    else:                                                  # model.py:372-392
        c = dense(drop(dense(m, "initialize/fc_a1", torch.tanh), im.get("init_a"), kf), "initialize/fc_a2")
        h_state = dense(drop(dense(m, "initialize/fc_b1", torch.tanh), im.get("init_b"), kf), "initialize/fc_b2")
    h_out = h_state
    word = torch.zeros(B, dtype=torch.long)
    ces, alphas, correct = [], [], []
    ctx2d = ctx.reshape(B * L, -1)
    for t in range(T):
        dm = step_masks(cfg, seed, t, B) if seed is not None else {}
        if cfg.num_attend_layers == 1:                     # model.py:401-414 (both layers bias-free)
            l1 = (drop(ctx2d, dm.get("att_ctx"), kf) @ w["attend/fc_a/kernel"]).reshape(B, L)
            l2 = drop(h_out, dm.get("att_out"), kf) @ w["attend/fc_b/kernel"]
            e = l1 + l2
        else:                                              # model.py:415-434
            t1 = dense(drop(ctx2d, dm.get("att_ctx"), kf), "attend/fc_1a", torch.tanh)
            t2 = dense(drop(h_out, dm.get("att_out"), kf), "attend/fc_1b", torch.tanh)
            temp = t1 + t2.repeat_interleave(L, dim=0)
            temp = drop(temp, dm.get("att_mid"), kf)
            e = (temp @ w["attend/fc_2/kernel"]).reshape(B, L)
        alpha = torch.softmax(e, dim=1)
        context = (ctx * alpha[:, :, None]).sum(dim=1)
        alphas.append(alpha * mk[:, t:t + 1])
        emb = w["word_embedding/weights"][word]
        x = drop(torch.cat([context, emb], dim=1), dm.get("lstm_in"), kl)
        g = torch.cat([x, h_state], dim=1) @ w["lstm/lstm_cell/kernel"] + w["lstm/lstm_cell/bias"]
        i, j, f, o = torch.chunk(g, 4, dim=1)
        c = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
        h_raw = torch.sigmoid(o) * torch.tanh(c)
        h_out = drop(h_raw, dm.get("lstm_out"), kl)
        h_state = drop(h_raw, dm.get("lstm_state"), kl)
        expanded = drop(torch.cat([h_out, context, emb], dim=1), dm.get("dec_in"), kf)
        if cfg.num_decode_layers == 1:                     # model.py:442-447
            logits = dense(expanded, "decode/fc")
        else:                                              # model.py:448-458
            td = drop(dense(expanded, "decode/fc_1", torch.tanh), dm.get("dec_mid"), kf)
            logits = dense(td, "decode/fc_2")
        ce = torch.logsumexp(logits, dim=1) - logits[torch.arange(B), sent[:, t]]
        ces.append(ce * mk[:, t])
        pred = logits.argmax(dim=1)
        correct.append(torch.where(pred == sent[:, t], mk[:, t], torch.zeros_like(mk[:, t])))
        word = sent[:, t]
    msum = mk.sum() if global_mask_sum is None else torch.tensor(float(global_mask_sum), dtype=f64)
    gb = B if global_batch is None else int(global_batch)
    ce_loss = torch.stack(ces, 1).sum() / msum
    att = torch.stack(alphas, 2).sum(dim=2)
    att_loss = cfg.attention_loss_factor * ((1.0 - att) ** 2).sum() / 2.0 / (gb * L)
    reg = sum(cfg.fc_kernel_regularizer_scale * (w[n] ** 2).sum() / 2.0 for n in R.regularized_names(w))
    acc = torch.stack(correct, 1).sum() / msum
    return dict(cross_entropy_loss=ce_loss, attention_loss=att_loss, reg_loss=reg, accuracy=acc,
                total_loss=ce_loss + att_loss + reg)


def loss_and_grads(cfg, weights_np, contexts, sentences, masks, seed=None, global_mask_sum=None,
                   global_batch=None, reg_in_grad=True):
    """Losses (floats) and d total_loss / d w for every trainable tensor (numpy float64).

    reg_in_grad=False leaves the regulariser's gradient out (a data-parallel shard adds it once, after the
    all-reduce)."""
    import torch
    w = {k: torch.tensor(np.asarray(v, np.float64), requires_grad=True) for k, v in weights_np.items()}
    out = forward_torch(cfg, w, contexts, sentences, masks, seed, global_mask_sum, global_batch)
    loss = out["total_loss"] if reg_in_grad else out["cross_entropy_loss"] + out["attention_loss"]
    grads = torch.autograd.grad(loss, [w[k] for k in w], allow_unused=True)
    g = {k: (np.zeros_like(weights_np[k], dtype=np.float64) if gi is None else gi.numpy()) for k, gi in zip(w, grads)}
    return {k: float(v) for k, v in out.items()}, g


def clip_and_adam(weights, grads, m, v, step, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-6, clip=5.0):
    """tf.contrib.layers.optimize_loss(clip_gradients=5.0) + tf.train.AdamOptimizer (model.py:479-511).

    clip_by_global_norm: g *= clip / max(norm, clip).  TF Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; w -= lr_t * m / (sqrt(v) + eps).  `step` counts from 1.
    Returns (new_w, new_m, new_v, global_norm)."""
    norm = float(np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values())))
    scale = clip / max(norm, clip)
    lr_t = lr * np.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    nw, nm, nv = {}, {}, {}
    for k in weights:
        g = grads[k].astype(np.float64) * scale
        nm[k] = beta1 * m[k] + (1 - beta1) * g
        nv[k] = beta2 * v[k] + (1 - beta2) * g * g
        nw[k] = weights[k].astype(np.float64) - lr_t * nm[k] / (np.sqrt(nv[k]) + eps)
    return nw, nm, nv, norm


def apply_optimizer(kind, weights, grads, slots, step, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-6, clip=5.0, decay=0.9,
                    momentum=0.0, centered=True, use_nesterov=True):
    """The four optimizers of model.py:479-503 behind optimize_loss(clip_gradients=clip): TF 1.x update rules in fp64.

    slots: list of dicts (name -> array) in the order Adam [m, v]; RMSProp [rms (initialised to ONES), mg, mom];
    Momentum [accumulator]; SGD [].  Returns (new_weights, new_slots, global_norm).
      RMSProp (training_ops ApplyCenteredRMSProp / ApplyRMSProp):
          ms = decay*ms + (1-decay)*g^2 ; mg = decay*mg + (1-decay)*g ; mom = momentum*mom + lr*g/sqrt(ms - mg^2 + eps) ; w -= mom
      Momentum (ApplyMomentum): acc = momentum*acc + g ; w -= lr*g + lr*momentum*acc (nesterov) | lr*acc
      SGD: w -= lr*g
    """
    if kind == "Adam":
        nw, nm, nv, norm = clip_and_adam(weights, grads, slots[0], slots[1], step, lr, beta1, beta2, eps, clip)
        return nw, [nm, nv], norm
    norm = float(np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values())))
    scale = clip / max(norm, clip)
    nw, ns = {}, [dict() for _ in slots]
    for k in weights:
        g = grads[k].astype(np.float64) * scale
        w = weights[k].astype(np.float64)
        if kind == "RMSProp":
            ms = decay * slots[0][k] + (1 - decay) * g * g
            denom = ms
            if centered:
                mg = decay * slots[1][k] + (1 - decay) * g
                ns[1][k] = mg
                denom = ms - mg * mg
            else:
                ns[1][k] = slots[1][k]
            mom = momentum * slots[2][k] + lr * g / np.sqrt(denom + eps)
            ns[0][k], ns[2][k] = ms, mom
            nw[k] = w - mom
        elif kind == "Momentum":
            acc = momentum * slots[0][k] + g
            ns[0][k] = acc
            nw[k] = w - (lr * g + lr * momentum * acc if use_nesterov else lr * acc)
        elif kind == "SGD":
            nw[k] = w - lr * g
        else:
            raise ValueError(kind)
    return nw, ns, norm
