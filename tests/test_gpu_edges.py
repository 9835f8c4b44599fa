"""Edge cases of the decode / beam / training entry points against the oracle: the smallest shapes, batches smaller
than the static graph batch, captions that complete at once, fully masked rows, ids on the vocabulary boundary."""
import numpy as np
import pytest

from _util import SMALL, assert_close, make_pair
from oracle import ref_step as R
from oracle import train_ref as TR

pytestmark = pytest.mark.gpu


def test_single_image_single_step():
    ocfg, w, m = make_pair(1, **SMALL)
    ctx = R.synth_contexts(ocfg, 1)
    toks_ref, steps = R.decode_loop(ocfg, w, ctx, 1, None, np.float64)
    toks, logits = m.decode_loop(ctx, 1, None, want_logits=True)
    assert toks.shape == (1, 1)
    assert_close(logits[0], steps[0]["logits"], "logits")
    assert toks[0, 0] == toks_ref[0, 0]
    c0, h0 = m.initialize(ctx)
    rc, rh = R.initialize(ocfg, w, ctx, np.float64)
    assert_close(c0, rc, "c0")
    assert_close(h0, rh, "h0")


@pytest.mark.parametrize("n", [1, 3, 7])
def test_fewer_images_than_the_static_batch(n):
    """The reference pads its last batch with fake images (dataset.py:51-54); the ABI takes the real count."""
    ocfg, w, m = make_pair(7, **SMALL)
    ctx = R.synth_contexts(ocfg, 7)[:n]
    sub = R.OracleConfig(batch_size=n, **SMALL)
    _, steps = R.decode_loop(sub, w, ctx, 4, None, np.float64)
    toks, logits = m.decode_loop(ctx, 4, None, want_logits=True)
    assert toks.shape == (n, 4)
    for t in range(4):
        assert_close(logits[t], steps[t]["logits"], "logits step %d" % t)


def test_beam_search_when_the_first_word_ends_the_caption():
    """eos = the most probable first word: the best beam completes at step 0 (base_model.py:229-232) while the other
    beams carry on; partial and complete captions are sorted as the reference does."""
    dims = dict(SMALL)
    dims["max_caption_length"] = 5
    ocfg, w, m = make_pair(4, beam=3, **dims)
    ctx = R.synth_contexts(ocfg, 4)
    first = R.beam_search(ocfg, w, ctx, eos_id=-1, dtype=np.float64)
    eos = int(first[0][0].sentence[0])
    ref = R.beam_search(ocfg, w, ctx, eos_id=eos, dtype=np.float64)
    got = m.beam_search(ctx, eos_id=eos)
    assert len(ref[0][0].sentence) == 1                       # really completed at once
    for k in range(4):
        assert len(got[k]) == len(ref[k])
        for g, e in zip(got[k], ref[k]):
            assert g.sentence == [int(x) for x in e.sentence], "image %d" % k
            assert abs(g.score - e.score) <= 1e-3 * abs(e.score)


TDIMS = dict(num_ctx=9, dim_ctx=64, dim_embedding=32, num_lstm_units=32, dim_initalize_layer=16,
             dim_attend_layer=24, dim_decode_layer=40, vocabulary_size=50, max_caption_length=5)


def test_training_with_a_fully_masked_row_and_boundary_ids():
    """A caption of length 0 (mask row all zeros) contributes nothing to the cross entropy or the coverage sum; ids
    0 and V-1 are ordinary words."""
    ocfg, w, m = make_pair(4, seed=3, **TDIMS)
    T, V = ocfg.max_caption_length, ocfg.vocabulary_size
    ctx = R.synth_contexts(ocfg, 4, 3)
    sent = np.array([[0, V - 1, 1, 2, 3], [V - 1, 0, V - 1, 0, 5], [7, 8, 9, 10, 11], [4, 4, 4, 4, 4]], np.int32)
    masks = np.array([[1, 1, 1, 0, 0], [0, 0, 0, 0, 0], [1, 1, 1, 1, 1], [1, 0, 0, 0, 0]], np.float32)
    m.train_setup(4, T, weights=w)
    for seed in (0, 9):
        ref_l, ref_g = TR.loss_and_grads(ocfg, w, ctx, sent, masks, seed if seed else None, reg_in_grad=False)
        ce, acc, att, reg = [float(x) for x in m.train_forward_backward(ctx, sent, masks, seed=seed).cpu().numpy()]
        assert abs(ce - ref_l["cross_entropy_loss"]) < 1e-4 * ref_l["cross_entropy_loss"]
        assert abs(att - ref_l["attention_loss"]) < 1e-4 * ref_l["attention_loss"] + 1e-9
        assert abs(acc - ref_l["accuracy"]) < 1e-6
        got = {k: v.detach().cpu().numpy() for k, v in m.train_state_dict("grads").items()}
        floor = 1e-3 * max(np.abs(g).max() for g in ref_g.values())
        for k, g in ref_g.items():
            err = np.abs(got[k].reshape(g.shape) - g).max() / max(np.abs(g).max(), floor)
            assert err < 2e-4, "%s: %.3e" % (k, err)
    assert m.info("train_bad_ids") == 0
