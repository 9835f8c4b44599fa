"""Properties the reference's arithmetic implies, checked on the oracle."""
import numpy as np
import pytest

from oracle import ref_step as R


def small_cfg(**kw):
    base = dict(num_ctx=49, dim_ctx=64, dim_embedding=32, num_lstm_units=64, dim_initalize_layer=32,
                dim_attend_layer=32, dim_decode_layer=64, vocabulary_size=300, batch_size=3,
                max_caption_length=6)
    base.update(kw)
    return R.OracleConfig(**base)


@pytest.mark.parametrize("layers", [1, 2])
def test_step_shapes_and_row_sums(layers):
    cfg = small_cfg(num_attend_layers=layers, num_decode_layers=layers, num_initalize_layers=layers)
    w = R.init_weights(cfg, 3)
    ctx = R.synth_contexts(cfg, 3, 3)
    c0, h0 = R.initialize(cfg, w, ctx)
    r = R.decode_step(cfg, w, ctx, np.array([0, 5, 7], np.int32), c0, h0)
    assert r["memory"].shape == (3, 64) and r["probs"].shape == (3, 300) and r["alpha"].shape == (3, 49)
    np.testing.assert_allclose(r["alpha"].sum(1), 1.0, rtol=1e-5)
    np.testing.assert_allclose(r["probs"].sum(1), 1.0, rtol=1e-5)
    assert r["memory"].dtype == np.float32


def test_n1_two_layer_attention_is_independent_of_hidden_state():
    cfg = small_cfg()
    w = R.init_weights(cfg, 4)
    ctx = R.synth_contexts(cfg, 3, 4)
    rng = np.random.RandomState(0)
    a1 = R.attend(cfg, w, ctx, rng.randn(3, 64), np.float64)
    a2 = R.attend(cfg, w, ctx, rng.randn(3, 64) * 5, np.float64)
    assert np.abs(a1 - a2).max() < 1e-12


def test_one_layer_attention_depends_on_hidden_state():
    cfg = small_cfg(num_attend_layers=1)
    w = R.init_weights(cfg, 4)
    ctx = R.synth_contexts(cfg, 3, 4)
    rng = np.random.RandomState(0)
    a1 = R.attend(cfg, w, ctx, rng.randn(3, 64), np.float64)
    a2 = R.attend(cfg, w, ctx, rng.randn(3, 64), np.float64)
    assert np.abs(a1 - a2).max() > 1e-4


def test_fp32_and_fp64_oracles_agree():
    cfg = small_cfg()
    w = R.init_weights(cfg, 5)
    ctx = R.synth_contexts(cfg, 3, 5)
    t32, s32 = R.decode_loop(cfg, w, ctx, 6, None, np.float32)
    t64, s64 = R.decode_loop(cfg, w, ctx, 6, None, np.float64)
    for a, b in zip(s32, s64):
        assert np.abs(a["logits"] - b["logits"]).max() < 1e-4 * np.abs(b["logits"]).max()


def test_dropout_formula_and_masks():
    x = np.arange(8, dtype=np.float64).reshape(2, 4)
    m = np.array([[1, 0, 1, 0], [0, 1, 1, 1]], np.float64)
    np.testing.assert_allclose(R._dropout(x, m, 0.5), x / 0.5 * m)
    assert R._dropout(x, None, 0.5) is x


def test_topn_matches_reference_heap_semantics():
    t = R.TopN(3)
    for s in [0.1, 0.5, 0.3, 0.5, 0.2, 0.9]:
        t.push(R.CaptionData([], None, None, s))
    got = [c.score for c in t.extract(sort=True)]
    assert got == [0.9, 0.5, 0.5]


def test_beam_search_oracle_invariants():
    cfg = small_cfg(beam_size=3)
    w = R.init_weights(cfg, 6)
    ctx = R.synth_contexts(cfg, 2, 6)
    res = R.beam_search(cfg, w, ctx, eos_id=2)
    assert len(res) == 2
    for caps in res:
        assert 1 <= len(caps) <= 3
        sc = [c.score for c in caps]
        assert sc == sorted(sc, reverse=True)
        for c in caps:
            assert 0 < c.score <= 1 and len(c.sentence) <= cfg.max_caption_length
    # beam 1 with an unreachable eos == greedy decoding
    cfg1 = small_cfg(beam_size=1)
    res1 = R.beam_search(cfg1, w, ctx, eos_id=-1)
    toks, _ = R.decode_loop(cfg1, w, ctx, cfg1.max_caption_length)
    for k in range(2):
        assert res1[k][0].sentence == list(toks[k])


def test_train_forward_dropout_off_equals_inference_steps():
    cfg = small_cfg()
    w = R.init_weights(cfg, 8)
    ctx = R.synth_contexts(cfg, 3, 8)
    rng = np.random.RandomState(1)
    sent = rng.randint(1, 300, (3, 6)).astype(np.int32)
    masks = np.ones((3, 6), np.float32)
    out = R.train_forward(cfg, w, ctx, sent, masks, np.float64)
    _, steps = R.decode_loop(cfg, w, ctx, 6, sent, np.float64)
    for a, b in zip(out["logits"], steps):
        np.testing.assert_allclose(a, b["logits"], rtol=1e-10, atol=1e-12)


def test_hoisted_stepper_and_fast_topk_equal_the_literal_restatement():
    """The two shortcuts the full-size beam-search parity test uses are exact: the hoisted fc_1a branch gives
    bit-identical steps, and the stable argsort picks the same words in the same order as the reference's list sort."""
    import numpy as np
    from oracle import ref_step as R
    cfg = R.OracleConfig(batch_size=3, beam_size=3, num_ctx=25, dim_ctx=32, dim_embedding=16, num_lstm_units=32,
                         dim_initalize_layer=16, dim_attend_layer=24, dim_decode_layer=32, vocabulary_size=40,
                         max_caption_length=6)
    w = R.init_weights(cfg, seed=4)
    ctx = R.synth_contexts(cfg, 3, seed=4)
    rng = np.random.RandomState(1)
    lw = rng.randint(0, 40, 3).astype(np.int32)
    c = rng.uniform(-0.5, 0.5, (3, 32)).astype(np.float32)
    h = rng.uniform(-0.5, 0.5, (3, 32)).astype(np.float32)
    for dt in (np.float32, np.float64):
        st = R.HoistedStepper(cfg, w, ctx, dt)
        a = R.decode_step(cfg, w, ctx, lw, c, h, dt)
        mem, out, probs = st.step(ctx, lw, c, h)
        assert np.array_equal(mem, a["memory"]) and np.array_equal(out, a["output"]) and np.array_equal(probs, a["probs"])
    # ties included: quantised probabilities produce many equal scores
    def coarse(cx, lw_, lm, lo):
        r = R.decode_step(cfg, w, cx, lw_, lm, lo, np.float64)
        return r["memory"], r["output"], np.round(r["probs"], 2)
    a = R.beam_search(cfg, w, ctx, eos_id=2, step_fn=coarse)
    b = R.beam_search(cfg, w, ctx, eos_id=2, step_fn=coarse, fast_topk=True)
    for x, y in zip(a, b):
        assert [c_.sentence for c_ in x] == [c_.sentence for c_ in y]
        assert [c_.score for c_ in x] == [c_.score for c_ in y]
