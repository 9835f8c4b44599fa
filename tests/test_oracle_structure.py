"""The oracle restates the graph TensorFlow actually built: check it against the GraphDef the
reference recorded in its own event file (tests/golden/graph_fixture.json, extracted by
tests/golden/make_graph_fixture.py)."""
import json
import os

import numpy as np

from oracle import ref_step as R

FIX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "graph_fixture.json")))


def test_variable_names_and_shapes_match_recorded_graph():
    cfg = R.OracleConfig()  # reference defaults: config.py:8-17, vgg16 196x512, vocab 5000
    mine = {k: list(v) for k, v in R.weight_shapes(cfg).items()}
    assert mine == FIX["variables"]
    assert sum(int(np.prod(v)) for v in mine.values()) == 13983112


def test_lstm_wiring_gate_order_and_forget_bias():
    w = FIX["wiring"]
    # split:0 -> sigmoid (i), split:1 -> tanh (j), split:2 + 1.0 -> sigmoid (f), split:3 -> sigmoid (o)
    assert w["lstm/lstm_cell/Sigmoid_1"]["inputs"] == ["lstm/lstm_cell/split"]
    assert w["lstm/lstm_cell/Tanh"]["inputs"] == ["lstm/lstm_cell/split:1"]
    assert w["lstm/lstm_cell/add"]["inputs"][0] == "lstm/lstm_cell/split:2"
    assert FIX["consts"]["lstm/lstm_cell/add/y"] == [1.0]
    assert w["lstm/lstm_cell/Sigmoid_2"]["inputs"] == ["lstm/lstm_cell/split:3"]
    # c = sigmoid(f+1)*c_prev + sigmoid(i)*tanh(j) ; c_prev of step 0 is initialize/fc_a2 (memory)
    assert w["lstm/lstm_cell/mul"]["inputs"] == ["lstm/lstm_cell/Sigmoid", "initialize/fc_a2/BiasAdd"]
    assert w["lstm/lstm_cell/mul_1"]["inputs"] == ["lstm/lstm_cell/Sigmoid_1", "lstm/lstm_cell/Tanh"]
    assert w["lstm/lstm_cell/mul_2"]["inputs"] == ["lstm/lstm_cell/Sigmoid_2", "lstm/lstm_cell/Tanh_1"]
    # the oracle's cell reproduces exactly that
    cfg = R.OracleConfig(dim_ctx=32, dim_embedding=8, num_lstm_units=32)
    H = 32
    rng = np.random.RandomState(0)
    wts = {"lstm/lstm_cell/kernel": rng.randn(32 + 8 + H, 4 * H), "lstm/lstm_cell/bias": rng.randn(4 * H)}
    x, c, h = rng.randn(2, 40), rng.randn(2, H), rng.randn(2, H)
    g = np.concatenate([x, h], 1) @ wts["lstm/lstm_cell/kernel"] + wts["lstm/lstm_cell/bias"]
    i, j, f, o = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
    sig = lambda v: 1 / (1 + np.exp(-v))
    c_exp = sig(f + 1.0) * c + sig(i) * np.tanh(j)
    h_exp = sig(o) * np.tanh(c_exp)
    c_got, h_got = R.lstm_cell(cfg, wts, x, c, h)
    np.testing.assert_allclose(c_got, c_exp, rtol=1e-12)
    np.testing.assert_allclose(h_got, h_exp, rtol=1e-12)


def test_concat_orders():
    w = FIX["wiring"]
    # LSTM input = [context, word_embed] then h (model.py:277 + LSTMCell concat)
    assert w["lstm/concat"]["inputs"][:2] == ["attend/Sum", "word_embedding_1/embedding_lookup"]
    assert w["lstm/lstm_cell/concat"]["inputs"][:2] == ["lstm/dropout/mul", "initialize/fc_b2/BiasAdd"]
    # decode input = [output, context, word_embed] (model.py:283-286)
    assert w["decode/concat"]["inputs"][:3] == ["lstm/dropout_2/mul", "attend/Sum",
                                                "word_embedding_1/embedding_lookup"]
    # context vector multiplies the UN-dropped contexts
    assert w["attend/mul"]["inputs"][0] == "Reshape"
    # tanh is applied to the two attention branches separately, then added (N1)
    assert w["attend/add"]["inputs"] == ["attend/fc_1a/Tanh", "attend/Reshape_1"]
    # step 1 consumes state-dropout h and raw c of step 0 (a6')
    assert w["lstm/lstm_cell/concat_1"]["inputs"][:2] == ["lstm_1/dropout/mul", "lstm/dropout_1/mul"]


def test_recorded_reg_loss_identifies_regularised_set():
    """reg_loss = 1e-4 * sum(w^2)/2 over embedding + dense kernels, NOT the LSTM kernel / biases."""
    cfg = R.OracleConfig()
    shapes = R.weight_shapes(cfg)
    names = R.regularized_names(shapes)
    n = sum(int(np.prod(shapes[k])) for k in names)
    assert n == 10826240
    expected = 1e-4 * n * (0.08 ** 2 / 3) / 2      # E[w^2] of U(-0.08, 0.08)
    assert abs(expected - FIX["scalars"]["metrics/reg_loss"]) / expected < 5e-3


def test_recorded_step1_losses_are_consistent_with_oracle_statistics():
    """Random init, B=20: CE ~ ln(5000) + small, total = CE + attention + reg."""
    s = FIX["scalars"]
    assert abs(s["metrics/total_loss"] - (s["metrics/cross_entropy_loss"] + s["metrics/attention_loss"]
                                           + s["metrics/reg_loss"])) < 1e-4
    cfg = R.OracleConfig(batch_size=6, max_caption_length=8)
    w = R.init_weights(cfg, seed=1, random_bias=False)
    ctx = R.synth_contexts(cfg, 6, seed=1)
    rng = np.random.RandomState(2)
    sent = rng.randint(1, cfg.vocabulary_size, (6, 8)).astype(np.int32)
    masks = (np.arange(8)[None, :] < rng.randint(4, 9, 6)[:, None]).astype(np.float32)
    out = R.train_forward(cfg, w, ctx, sent, masks)
    assert abs(out["reg_loss"] - s["metrics/reg_loss"]) / s["metrics/reg_loss"] < 0.01
    assert abs(out["cross_entropy_loss"] - np.log(5000)) < 0.5
    assert 0 <= out["attention_loss"] < 0.05
