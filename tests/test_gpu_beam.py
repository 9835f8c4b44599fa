"""Device-side beam search against the oracle's restatement of base_model.py:163-240."""
import numpy as np
import pytest

from _util import SMALL, make_pair
from oracle import ref_step as R

pytestmark = pytest.mark.gpu


def compare(res, ref, tol=1e-3):
    assert len(res) == len(ref)
    for k, (got, exp) in enumerate(zip(res, ref)):
        assert len(got) == len(exp), "image %d: %d vs %d captions" % (k, len(got), len(exp))
        for g, e in zip(got, exp):
            assert g.sentence == [int(x) for x in e.sentence], "image %d" % k
            assert abs(g.score - e.score) <= tol * abs(e.score)


def pick_eos(ocfg, w, ctx):
    """an id that really occurs inside beams, so that captions complete at different steps"""
    res = R.beam_search(ocfg, w, ctx, eos_id=-1)
    words = [wd for caps in res for c in caps for wd in c.sentence[1:]]
    vals, counts = np.unique(words, return_counts=True)
    return int(vals[np.argmax(counts)])


@pytest.mark.parametrize("beam", [1, 3, 4])
def test_beam_search_small(beam):
    dims = dict(SMALL)
    dims["max_caption_length"] = 8
    ocfg, w, m = make_pair(5, beam=beam, **dims)
    ctx = R.synth_contexts(ocfg, 5)
    ref = R.beam_search(ocfg, w, ctx, eos_id=-1, dtype=np.float64)      # nothing ever completes
    compare(m.beam_search(ctx, eos_id=-1), ref)
    eos = pick_eos(ocfg, w, ctx)
    ref = R.beam_search(ocfg, w, ctx, eos_id=eos, dtype=np.float64)
    got = m.beam_search(ctx, eos_id=eos)
    compare(got, ref)
    assert any(c.complete for caps in got for c in caps)
    # device tensors in -> same answer, and a replayed graph gives it again
    import torch
    ctx_d = torch.from_numpy(ctx).cuda()
    for _ in range(3):
        compare(m.beam_search(ctx_d, eos_id=eos), ref)


def test_beam_search_reference_shapes():
    """beam=3 on the reference default graph (L=196, D=512, H=512, V=5000), 4 images."""
    ocfg, w, m = make_pair(4, beam=3, max_caption_length=6)
    ctx = R.synth_contexts(ocfg, 4)
    eos = pick_eos(ocfg, w, ctx)
    ref = R.beam_search(ocfg, w, ctx, eos_id=eos, dtype=np.float32)
    compare(m.beam_search(ctx, eos_id=eos), ref)


def test_beam_one_equals_greedy_loop():
    dims = dict(SMALL)
    ocfg, w, m = make_pair(6, beam=1, **dims)
    ctx = R.synth_contexts(ocfg, 6)
    toks = m.decode_loop(ctx, 6)
    res = m.beam_search(ctx, eos_id=-1)
    for k in range(6):
        assert res[k][0].sentence == [int(x) for x in toks[k]]


def test_config5_as_stated():
    """BASELINE config 5 as stated: 128 images x beam 3, T=30, L=196, D=512, H=1024, V=10000 (bench.py --workload 5).

    With random weights the word distribution is almost flat, so over 128 x 3 x 30 top-(beam+1) selections a few
    decisions are numerical ties (relative gap below the GPU path's ~5e-6 error).  The oracle is therefore run twice —
    exact fp64, and fp64 with every probability perturbed by 3e-5 relative noise — and an image counts as decidable
    when both runs give the same captions; decidable images (>= 90 %) must match the GPU exactly, scores to 1e-3."""
    ocfg, w, m = make_pair(128, beam=3, num_lstm_units=1024, vocabulary_size=10000, max_caption_length=30)
    ctx = R.synth_contexts(ocfg, 128)
    stepper = R.HoistedStepper(ocfg, w, ctx, np.float64)
    eos = pick_eos_fast(ocfg, w, ctx[:8])
    ref = R.beam_search(ocfg, w, ctx, eos_id=eos, dtype=np.float64, step_fn=stepper.step, fast_topk=True)
    rng = np.random.RandomState(5)

    def noisy(cx, lw, lm, lo):
        mem, out, probs = stepper.step(cx, lw, lm, lo)
        return mem, out, probs * (1.0 + 3e-5 * rng.standard_normal(probs.shape))
    ref2 = R.beam_search(ocfg, w, ctx, eos_id=eos, dtype=np.float64, step_fn=noisy, fast_topk=True)
    got = m.beam_search(ctx, eos_id=eos)
    decidable = 0
    for k in range(128):
        same = len(ref[k]) == len(ref2[k]) and all(a.sentence == b.sentence for a, b in zip(ref[k], ref2[k]))
        if not same:
            continue
        decidable += 1
        compare([got[k]], [ref[k]])
    assert decidable >= 115, decidable
    assert any(c.complete for caps in got for c in caps)


def pick_eos_fast(ocfg, w, ctx):
    import dataclasses
    sub = dataclasses.replace(ocfg, max_caption_length=6)
    return pick_eos(sub, w, ctx)
