"""CPU oracle for the soft-attention LSTM decode path of show-attend-and-tell.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker / the timed CPU
baseline.  The product path (``sat_b200``) never imports this module.

PARITY UNPINNED: the reference is a TensorFlow-1.x graph program; TensorFlow is
not installable in this environment (no wheel, no network) and the reference
ships no tests, golden vectors or known-answer files for this path (SURVEY.md
§4, §8c).  This restatement is therefore pinned only *structurally*, against
the GraphDef the reference itself recorded in
``summary/events.out.tfevents.1535421942.CLARK-CL-LI`` (variable names/shapes,
concat orders, LSTM gate order and forget bias, dropout formula) — see
``tests/golden/graph_fixture.json`` and ``tests/test_oracle_structure.py`` — and
*numerically* only through the scalars of the one training step that file
recorded (loss identities and brackets, ``tests/test_oracle_recorded_step.py``):
no tensor of the reference is available to compare with.

Every function cites the reference lines it restates (paths are relative to
``/root/reference``).  Arithmetic is numpy; ``dtype`` selects fp32 (the
reference's precision, model.py:205-213) or fp64 ("truth" used to bound both
fp32 sides).  Weights are a dict keyed by the TF variable names without the
``:0`` suffix (base_model.py:242-278 save/load format).
"""
from __future__ import annotations

import heapq
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------
# configuration (config.py:4-44) — field names kept, including the
# `initalize` typo, so parity tests read like the reference.
# --------------------------------------------------------------------------
@dataclass
class OracleConfig:
    num_ctx: int = 196                 # model.py:54-59 (vgg16 conv5_3 -> 196 locations)
    dim_ctx: int = 512
    max_caption_length: int = 20       # config.py:9
    dim_embedding: int = 512           # config.py:10
    num_lstm_units: int = 512          # config.py:11
    num_initalize_layers: int = 2      # config.py:12
    dim_initalize_layer: int = 512     # config.py:13
    num_attend_layers: int = 2         # config.py:14
    dim_attend_layer: int = 512        # config.py:15
    num_decode_layers: int = 2         # config.py:16
    dim_decode_layer: int = 1024       # config.py:17
    vocabulary_size: int = 5000        # config.py:67
    batch_size: int = 4
    beam_size: int = 3                 # main.py:35
    fc_kernel_initializer_scale: float = 0.08   # config.py:20
    fc_kernel_regularizer_scale: float = 1e-4   # config.py:21
    fc_drop_rate: float = 0.5          # config.py:25
    lstm_drop_rate: float = 0.3        # config.py:26
    attention_loss_factor: float = 0.01  # config.py:27


def weight_shapes(cfg: OracleConfig) -> Dict[str, Tuple[int, ...]]:
    """Variable names and shapes exactly as TF creates them for the decoder.

    Names: tf.layers.dense -> '<scope>/<name>/kernel' [in, units] and
    '<scope>/<name>/bias' [units] (utils/nn.py:85-105); the LSTM cell ->
    'lstm/lstm_cell/kernel' [D+E+H, 4H], 'lstm/lstm_cell/bias' [4H]
    (model.py:228-230, 276-279); embedding 'word_embedding/weights' [V,E]
    (model.py:219-225).  Cross-checked against the recorded GraphDef
    (tests/golden/graph_fixture.json).
    """
    D, E, H, V = cfg.dim_ctx, cfg.dim_embedding, cfg.num_lstm_units, cfg.vocabulary_size
    A, Dd, I, L = cfg.dim_attend_layer, cfg.dim_decode_layer, cfg.dim_initalize_layer, cfg.num_ctx
    s: Dict[str, Tuple[int, ...]] = {"word_embedding/weights": (V, E)}
    if cfg.num_initalize_layers == 1:            # model.py:362-371
        s["initialize/fc_a/kernel"] = (D, H); s["initialize/fc_a/bias"] = (H,)
        s["initialize/fc_b/kernel"] = (D, H); s["initialize/fc_b/bias"] = (H,)
    else:                                        # model.py:372-392
        for n in ("a", "b"):
            s[f"initialize/fc_{n}1/kernel"] = (D, I); s[f"initialize/fc_{n}1/bias"] = (I,)
            s[f"initialize/fc_{n}2/kernel"] = (I, H); s[f"initialize/fc_{n}2/bias"] = (H,)
    if cfg.num_attend_layers == 1:               # model.py:401-414 (both bias-free)
        s["attend/fc_a/kernel"] = (D, 1)
        s["attend/fc_b/kernel"] = (H, L)
    else:                                        # model.py:415-434
        s["attend/fc_1a/kernel"] = (D, A); s["attend/fc_1a/bias"] = (A,)
        s["attend/fc_1b/kernel"] = (H, A); s["attend/fc_1b/bias"] = (A,)
        s["attend/fc_2/kernel"] = (A, 1)
    s["lstm/lstm_cell/kernel"] = (D + E + H, 4 * H)
    s["lstm/lstm_cell/bias"] = (4 * H,)
    if cfg.num_decode_layers == 1:               # model.py:442-447
        s["decode/fc/kernel"] = (H + D + E, V); s["decode/fc/bias"] = (V,)
    else:                                        # model.py:448-458
        s["decode/fc_1/kernel"] = (H + D + E, Dd); s["decode/fc_1/bias"] = (Dd,)
        s["decode/fc_2/kernel"] = (Dd, V); s["decode/fc_2/bias"] = (V,)
    return s


def init_weights(cfg: OracleConfig, seed: int = 1234, random_bias: bool = True
                 ) -> Dict[str, np.ndarray]:
    """U(-0.08, 0.08) kernels (utils/nn.py:29-31, model.py:223,230).

    The reference initialises every bias to 0; parity runs use random biases
     so that bias plumbing is exercised.
    """
    rng = np.random.RandomState(seed)
    sc = cfg.fc_kernel_initializer_scale
    w = {}
    for name, shp in weight_shapes(cfg).items():
        if name.endswith("bias") and not random_bias:
            w[name] = np.zeros(shp, np.float32)
        else:
            w[name] = rng.uniform(-sc, sc, size=shp).astype(np.float32)
    return w


def synth_contexts(cfg: OracleConfig, batch: int, seed: int = 1234) -> np.ndarray:
    """relu(N(0,1)) features: conv5_3 / res5c are ReLU outputs (model.py:52, 187)."""
    rng = np.random.RandomState(seed + 1)
    return np.maximum(rng.standard_normal((batch, cfg.num_ctx, cfg.dim_ctx)), 0).astype(np.float32)


# --------------------------------------------------------------------------
# primitive layers
# --------------------------------------------------------------------------
def _dense(x, w, name, act=None, use_bias=True):
    """tf.layers.dense: y = act(x @ kernel + bias), kernel [in, units] (utils/nn.py:85-105)."""
    y = x @ w[name + "/kernel"].astype(x.dtype)
    if use_bias:
        y = y + w[name + "/bias"].astype(x.dtype)
    if act is not None:
        y = act(y)
    return y


def _dropout(x, mask, keep):
    """tf.layers.dropout / DropoutWrapper: x / keep * floor(keep + U[0,1)).

    ``mask`` is the injected 0/1 array ``floor(keep + U)``; ``None`` = inference.
    """
    if mask is None:
        return x
    return x / x.dtype.type(keep) * mask.astype(x.dtype)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _softmax(x):
    """tf.nn.softmax over the last axis (model.py:288, 435)."""
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=-1, keepdims=True)


# --------------------------------------------------------------------------
# model.py:358-393  initialize
# --------------------------------------------------------------------------
def initialize(cfg: OracleConfig, w, contexts, dtype=np.float32, masks: Optional[dict] = None):
    """(c0, h0) from the mean context.  model.py:239-242, 358-393.

    Returns (initial_memory, initial_output) = (c0, h0).
    """
    masks = masks or {}
    keep = 1.0 - cfg.fc_drop_rate
    ctx = contexts.astype(dtype)
    m = ctx.mean(axis=1)                                     # model.py:240
    m = _dropout(m, masks.get("init_mean"), keep)            # model.py:361
    if cfg.num_initalize_layers == 1:
        memory = _dense(m, w, "initialize/fc_a")             # model.py:364-367
        output = _dense(m, w, "initialize/fc_b")             # model.py:368-371
    else:
        t1 = _dense(m, w, "initialize/fc_a1", np.tanh)       # model.py:374-377
        t1 = _dropout(t1, masks.get("init_a"), keep)
        memory = _dense(t1, w, "initialize/fc_a2")           # model.py:379-382
        t2 = _dense(m, w, "initialize/fc_b1", np.tanh)       # model.py:384-387
        t2 = _dropout(t2, masks.get("init_b"), keep)
        output = _dense(t2, w, "initialize/fc_b2")           # model.py:389-392
    return memory, output


# --------------------------------------------------------------------------
# model.py:395-436  attend
# --------------------------------------------------------------------------
def attend(cfg: OracleConfig, w, contexts, output, dtype=np.float32, masks: Optional[dict] = None, t1=None):
    """alpha [B, L] = softmax over locations.  model.py:395-436.

    t1 (2-layer inference only): the fc_1a branch tanh(ctx2d @ W1a + b1a) computed by an earlier call on the same
    contexts (it does not depend on the state: the reference recomputes it every step, the result is the same array)."""
    masks = masks or {}
    keep = 1.0 - cfg.fc_drop_rate
    B, L, D = contexts.shape
    ctx2d = contexts.astype(dtype).reshape(B * L, D)                     # model.py:398
    ctx2d = _dropout(ctx2d, masks.get("att_ctx"), keep)                  # model.py:399
    out = _dropout(output.astype(dtype), masks.get("att_out"), keep)     # model.py:400
    if cfg.num_attend_layers == 1:
        l1 = _dense(ctx2d, w, "attend/fc_a", None, use_bias=False).reshape(B, L)   # model.py:403-408
        l2 = _dense(out, w, "attend/fc_b", None, use_bias=False)                     # model.py:409-413
        logits = l1 + l2                                                             # model.py:414
    else:
        if t1 is None:
            t1 = _dense(ctx2d, w, "attend/fc_1a", np.tanh)               # model.py:417-420
        t2 = _dense(out, w, "attend/fc_1b", np.tanh)                     # model.py:421-424
        t2 = np.repeat(t2[:, None, :], L, axis=1).reshape(B * L, -1)     # model.py:425-426 (tile)
        t = t1 + t2                                                      # model.py:427
        t = _dropout(t, masks.get("att_mid"), keep)                      # model.py:428
        logits = _dense(t, w, "attend/fc_2", None, use_bias=False).reshape(B, L)     # model.py:429-434
    return _softmax(logits)                                              # model.py:435


# --------------------------------------------------------------------------
# TF-1.7 LSTMCell (un-vendored dependency; semantics confirmed on the GraphDef:
# concat([x, h]) @ kernel + bias, split into i, j, f, o, forget_bias 1.0)
# --------------------------------------------------------------------------
def lstm_cell(cfg: OracleConfig, w, x, c_prev, h_prev):
    """One LSTMCell call (model.py:228-230, 278).  Returns (c, h).

    Graph fixture nodes: lstm/lstm_cell/{concat, MatMul, BiasAdd, split,
    add (y=1.0), Sigmoid, Sigmoid_1, Tanh, mul, mul_1, add_1, Sigmoid_2,
    Tanh_1, mul_2}.
    """
    dt = x.dtype
    g = np.concatenate([x, h_prev], axis=1) @ w["lstm/lstm_cell/kernel"].astype(dt) \
        + w["lstm/lstm_cell/bias"].astype(dt)
    i, j, f, o = np.split(g, 4, axis=1)
    c = _sigmoid(f + dt.type(1.0)) * c_prev + _sigmoid(i) * np.tanh(j)
    h = _sigmoid(o) * np.tanh(c)
    return c, h


# --------------------------------------------------------------------------
# model.py:438-459  decode
# --------------------------------------------------------------------------
def decode(cfg: OracleConfig, w, expanded_output, masks: Optional[dict] = None):
    """Word logits [B, V].  model.py:438-459."""
    masks = masks or {}
    keep = 1.0 - cfg.fc_drop_rate
    x = _dropout(expanded_output, masks.get("dec_in"), keep)             # model.py:441
    if cfg.num_decode_layers == 1:
        return _dense(x, w, "decode/fc")                                 # model.py:444-447
    t = _dense(x, w, "decode/fc_1", np.tanh)                             # model.py:450-453
    t = _dropout(t, masks.get("dec_mid"), keep)                          # model.py:454
    return _dense(t, w, "decode/fc_2")                                   # model.py:455-458


# --------------------------------------------------------------------------
# model.py:258-290  one decode step (inference graph: num_steps = 1, model.py:255)
# --------------------------------------------------------------------------
def decode_step(cfg: OracleConfig, w, contexts, last_word, last_memory, last_output,
                dtype=np.float32, t1=None):
    """One inference step: the sess.run of base_model.py:207-212.

    feeds: contexts [B,L,D] f32, last_word [B] i32, last_memory (=c) [B,H],
    last_output (=h) [B,H].  Returns dict(memory, output, probs, logits, alpha,
    context).  `memory, _ = state` (model.py:279): memory = c, output = h.
    """
    ctx = contexts.astype(dtype)
    c_prev = last_memory.astype(dtype)
    h_prev = last_output.astype(dtype)
    alpha = attend(cfg, w, ctx, h_prev, dtype, None, t1)                  # model.py:262
    context = (ctx * alpha[:, :, None]).sum(axis=1)                       # model.py:263-264
    word_embed = w["word_embedding/weights"].astype(dtype)[np.asarray(last_word)]  # model.py:273
    current_input = np.concatenate([context, word_embed], axis=1)         # model.py:277
    memory, output = lstm_cell(cfg, w, current_input, c_prev, h_prev)     # model.py:278-279
    expanded = np.concatenate([output, context, word_embed], axis=1)      # model.py:283-286
    logits = decode(cfg, w, expanded)                                     # model.py:287
    probs = _softmax(logits)                                              # model.py:288
    return dict(memory=memory, output=output, probs=probs, logits=logits,
                alpha=alpha, context=context)


def decode_loop(cfg: OracleConfig, w, contexts, num_steps: int,
                forced_words: Optional[np.ndarray] = None, dtype=np.float32):
    """T steps from initialize(); greedy (argmax, model.py:289) or teacher-forced.

    Word fed at step 0 is 0 (<start>; model.py:254, base_model.py:193-194).
    forced_words [B,T]: word fed at step t+1 is forced_words[:, t] (model.py:310).
    Returns tokens [B,T] (argmax per step) and the list of per-step dicts.
    """
    B = contexts.shape[0]
    c, h = initialize(cfg, w, contexts, dtype)
    word = np.zeros(B, np.int32)
    toks, steps = [], []
    for t in range(num_steps):
        r = decode_step(cfg, w, contexts, word, c, h, dtype)
        c, h = r["memory"], r["output"]
        pred = r["logits"].argmax(axis=1).astype(np.int32)               # model.py:289
        toks.append(pred); steps.append(r)
        word = forced_words[:, t].astype(np.int32) if forced_words is not None else pred
    return np.stack(toks, axis=1), steps


class HoistedStepper:
    """decode_step on fixed contexts with the state-independent fc_1a branch of `attend` computed once (the SAME numpy
    expression the step would evaluate, so the results are bit-identical to decode_step's): makes full-size beam
    searches (128 images x 3 beams x 30 steps) affordable for the parity tests."""

    def __init__(self, cfg: OracleConfig, w, contexts, dtype=np.float32):
        self.cfg, self.w, self.dtype = cfg, w, dtype
        self.ctx = contexts.astype(dtype)
        B, L, D = contexts.shape
        self.t1 = (_dense(self.ctx.reshape(B * L, D), w, "attend/fc_1a", np.tanh)
                   if cfg.num_attend_layers == 2 else None)

    def step(self, contexts, last_word, last_memory, last_output):
        r = decode_step(self.cfg, self.w, self.ctx, last_word, last_memory, last_output, self.dtype, self.t1)
        return r["memory"], r["output"], r["probs"]


# --------------------------------------------------------------------------
# utils/misc.py:38-87  CaptionData / TopN ; base_model.py:163-240 beam_search
# --------------------------------------------------------------------------
class CaptionData:
    """utils/misc.py:38-60: ordering is by score only."""
    __slots__ = ("sentence", "memory", "output", "score")

    def __init__(self, sentence, memory, output, score):
        self.sentence, self.memory, self.output, self.score = sentence, memory, output, score

    def __lt__(self, other):
        return self.score < other.score

    def __eq__(self, other):
        return self.score == other.score


class TopN:
    """utils/misc.py:62-87: size-n min-heap that keeps the n largest."""

    def __init__(self, n):
        self._n, self._data = n, []

    def size(self):
        return len(self._data)

    def push(self, x):
        if len(self._data) < self._n:
            heapq.heappush(self._data, x)
        else:
            heapq.heappushpop(self._data, x)

    def extract(self, sort=False):
        data, self._data = self._data, None
        if sort:
            data.sort(reverse=True)
        return data

    def reset(self):
        self._data = []


def beam_search(cfg: OracleConfig, w, contexts, eos_id: int, dtype=np.float32,
                step_fn=None, fast_topk=False):
    """base_model.py:163-240 with precomputed contexts in place of images.

    `vocabulary.words[w] == '.'` (base_model.py:229) is `w == eos_id`.
    Scores are products of probabilities (base_model.py:224) accumulated in
    Python floats (fp64), exactly as the reference does with numpy scalars.
    Returns, per image, the list of CaptionData sorted by descending score.
    `step_fn(contexts, last_word, last_memory, last_output) -> (memory, output,
    probs)` lets the tests drive this host loop with the CUDA step.
    fast_topk: take the beam+1 best words with a stable numpy argsort instead of sorting the
    Python list of (word, score) pairs (base_model.py:216-219) — the same order, ties included
    (a stable sort on -score keeps equal scores in index order, like list.sort with that key).
    """
    B = contexts.shape[0]
    if step_fn is None:
        def step_fn(ctx, lw, lm, lo):
            r = decode_step(cfg, w, ctx, lw, lm, lo, dtype)
            return r["memory"], r["output"], r["probs"]
    initial_memory, initial_output = initialize(cfg, w, contexts, dtype)   # base_model.py:168-170
    partial, complete = [], []
    for k in range(B):                                                     # base_model.py:174-181
        partial.append(TopN(cfg.beam_size))
        partial[-1].push(CaptionData([], initial_memory[k], initial_output[k], 1.0))
        complete.append(TopN(cfg.beam_size))
    for idx in range(cfg.max_caption_length):                              # base_model.py:184
        lists = []
        for k in range(B):
            lists.append(partial[k].extract()); partial[k].reset()
        num_steps = 1 if idx == 0 else cfg.beam_size                       # base_model.py:191
        for b in range(num_steps):
            if idx == 0:
                last_word = np.zeros(B, np.int32)                          # base_model.py:193-194
            else:
                last_word = np.array([pcl[b].sentence[-1] for pcl in lists], np.int32)
            last_memory = np.array([pcl[b].memory for pcl in lists], np.float32)
            last_output = np.array([pcl[b].output for pcl in lists], np.float32)
            memory, output, scores = step_fn(contexts, last_word, last_memory, last_output)
            for k in range(B):                                             # base_model.py:215-232
                cd = lists[k][b]
                if fast_topk:
                    order = np.argsort(-np.asarray(scores[k]), kind="stable")[:cfg.beam_size + 1]
                    ws = [(int(i), scores[k][i]) for i in order]
                else:
                    ws = list(enumerate(scores[k]))
                    ws.sort(key=lambda x: -x[1])
                for wd, s in ws[:cfg.beam_size + 1]:
                    beam = CaptionData(cd.sentence + [wd], memory[k], output[k],
                                       float(cd.score) * float(s))
                    if wd == eos_id:
                        complete[k].push(beam)
                    else:
                        partial[k].push(beam)
    results = []
    for k in range(B):                                                     # base_model.py:234-238
        if complete[k].size() == 0:
            complete[k] = partial[k]
        results.append(complete[k].extract(sort=True))
    return results


# --------------------------------------------------------------------------
# utils/vocabulary.py:53-63  Vocabulary.get_sentence
# --------------------------------------------------------------------------
def get_sentence(words: Sequence[str], idxs: Sequence[int]) -> str:
    """Word ids -> caption text, utils/vocabulary.py:53-63: append '.' unless the last word is one, cut after
    the first '.', join with a space before every token that is neither punctuation nor starts with "'"."""
    import string
    ws = [words[i] for i in idxs]
    if ws[-1] != '.':
        ws.append('.')
    length = int(np.argmax(np.array(ws) == '.')) + 1
    ws = ws[:length]
    return "".join([" " + w if not w.startswith("'") and w not in string.punctuation else w for w in ws]).strip()


# --------------------------------------------------------------------------
# model.py:250-334  training forward (teacher forcing, injected dropout masks)
# --------------------------------------------------------------------------
def regularized_names(w) -> List[str]:
    """L2-regularised set: embedding + dense kernels; NOT the LSTM kernel, NOT
    biases (utils/nn.py:33-37, model.py:224; verified by the recorded
    reg_loss = 1.154333)."""
    return [n for n in w if (n.endswith("/kernel") and not n.startswith("lstm/"))
            or n == "word_embedding/weights"]


def train_forward(cfg: OracleConfig, w, contexts, sentences, masks, dtype=np.float64,
                  drop_masks: Optional[List[dict]] = None, init_masks: Optional[dict] = None):
    """Forward pass of the unrolled training graph, model.py:250-334.

    drop_masks[t] may hold the per-step injected 0/1 masks: att_ctx, att_out,
    att_mid, dec_in, dec_mid (fc dropout, keep 0.5) and lstm_in [B,D+E],
    lstm_state [B,H], lstm_out [B,H] (DropoutWrapper, keep 0.7; state dropout
    applies to h only — SURVEY.md a6').  None = all dropout off.
    Returns dict of losses and per-step alphas/logits.
    """
    B, T = sentences.shape
    L = cfg.num_ctx
    ctx = contexts.astype(dtype)
    mk = masks.astype(dtype)
    keep_l = 1.0 - cfg.lstm_drop_rate
    c, h_state = initialize(cfg, w, ctx, dtype, init_masks)
    h_out = h_state                          # last_output = initial_output (model.py:251)
    word = np.zeros(B, np.int32)             # model.py:254
    ces, alphas, correct, logits_all = [], [], [], []
    for t in range(T):
        dm = drop_masks[t] if drop_masks is not None else {}
        alpha = attend(cfg, w, ctx, h_out, dtype, dm)                     # model.py:262
        context = (ctx * alpha[:, :, None]).sum(axis=1)                   # un-dropped ctx (a4)
        alphas.append(alpha * mk[:, t:t + 1])                             # model.py:266-269
        emb = w["word_embedding/weights"].astype(dtype)[word]
        x = np.concatenate([context, emb], axis=1)
        x = _dropout(x, dm.get("lstm_in"), keep_l)                        # DropoutWrapper input
        c, h_raw = lstm_cell(cfg, w, x, c, h_state)
        h_out = _dropout(h_raw, dm.get("lstm_out"), keep_l)               # output dropout
        h_state = _dropout(h_raw, dm.get("lstm_state"), keep_l)           # state dropout (h only)
        expanded = np.concatenate([h_out, context, emb], axis=1)          # model.py:283-286
        logits = decode(cfg, w, expanded, dm)
        logits_all.append(logits)
        lse = np.log(np.exp(logits - logits.max(1, keepdims=True)).sum(1)) + logits.max(1)
        ce = lse - logits[np.arange(B), sentences[:, t]]                  # model.py:294-296
        ces.append(ce * mk[:, t])                                         # model.py:297
        pred = logits.argmax(1)
        correct.append(np.where(pred == sentences[:, t], mk[:, t], 0.0))  # model.py:300-305
        word = sentences[:, t].astype(np.int32)                           # model.py:310
    ce_loss = np.stack(ces, 1).sum() / mk.sum()                           # model.py:316-318
    att = np.stack(alphas, 2).sum(axis=2)                                 # model.py:320-322 ([B,L])
    diffs = 1.0 - att
    att_loss = cfg.attention_loss_factor * (diffs ** 2).sum() / 2.0 / (B * L)   # model.py:323-326
    reg = sum(cfg.fc_kernel_regularizer_scale * (w[n].astype(dtype) ** 2).sum() / 2.0
              for n in regularized_names(w))                              # model.py:328
    acc = np.stack(correct, 1).sum() / mk.sum()                           # model.py:332-334
    return dict(total_loss=ce_loss + att_loss + reg, cross_entropy_loss=ce_loss,
                attention_loss=att_loss, reg_loss=reg, accuracy=acc,
                attentions=att, logits=logits_all)
