"""The one training step the reference itself recorded (summary/events.out.tfevents.1535421942.CLARK-CL-LI, step 1,
B=20, default graph, dropout on, random init) against the training oracle (oracle/train_ref.py).

These are the only reference-held NUMBERS for the path: metrics/{cross_entropy,attention,reg}_loss,
attentions/{mean,stddev,min,max} and optimizer/.../gradient_norm.  The inputs of that step (20 COCO images, their
captions, TF's unseeded masks and initial weights) are not recorded, so the pin is (a) exact identities between the
recorded scalars that the oracle's loss formulas must satisfy, and (b) brackets: the oracle run at the recorded setting
with the reference's own caption data must land around the recorded cross entropy and gradient norm.
tests/golden/vgg_conv5_3.npz (made by tests/golden/make_vgg_features.py from the files the reference ships) supplies
real captions and real conv5_3 features.
"""
import json
import os
import warnings

import numpy as np

from oracle import ref_step as R
from oracle import train_ref as TR

GOLD = os.path.join(os.path.dirname(__file__), "golden")
S = json.load(open(os.path.join(GOLD, "graph_fixture.json")))["scalars"]
REC_CE, REC_ATT, REC_NORM = (S["metrics/cross_entropy_loss"], S["metrics/attention_loss"],
                             S["optimizer/OptimizeLoss/global_norm/gradient_norm"])


def _fixture():
    z = np.load(os.path.join(GOLD, "vgg_conv5_3.npz"))
    return z["feats"].astype(np.float32), z["sentences"], z["masks"], json.loads(str(z["stats"]))


def test_recorded_attention_loss_is_the_oracles_formula_of_the_recorded_attention_statistics():
    """model.py:320-326: attention_loss = factor * l2_loss(1 - attentions) / (B*L), l2_loss = sum(x^2)/2, i.e.
    0.01/2 * mean((1-att)^2) = 0.005 * ((1 - mean)^2 + stddev^2): the recorded scalars satisfy it to 1e-6."""
    mean, std = S["attentions/mean"], S["attentions/stddev"]
    assert abs(0.005 * ((1.0 - mean) ** 2 + std ** 2) - REC_ATT) < 2e-6 * REC_ATT + 1e-9
    # ... and the oracle's loss is that formula: any attention map with these two moments gives the recorded loss
    cfg = R.OracleConfig(batch_size=20)
    rng = np.random.RandomState(0)
    att = rng.standard_normal((20, cfg.num_ctx))
    att = (att - att.mean()) / att.std() * std + mean
    diffs = 1.0 - att
    loss = cfg.attention_loss_factor * (diffs ** 2).sum() / 2.0 / (20 * cfg.num_ctx)      # ref_step.train_forward
    assert abs(loss - REC_ATT) < 2e-6 * REC_ATT


def test_recorded_attention_mean_is_the_mean_caption_length_over_L():
    """attentions = sum_t alpha_t * mask_t (model.py:266-269, 320-322) and every alpha_t sums to 1 over L, so
    mean(attentions) = mean caption length / L: 11.0 words for the recorded batch; the reference's own captions
    (<= 20 tokens, data/train/captions_train2014.json through data/vocabulary.csv) average 11.0 too."""
    _, sent, masks, stats = _fixture()
    assert abs(S["attentions/mean"] * 196 - 11.0) < 0.01
    assert abs(stats["mean_caption_length"] - S["attentions/mean"] * 196) < 0.5
    assert int(sent.max()) < 5000 and (masks.sum(1) >= 1).all()
    # the oracle reproduces the identity on a real batch
    cfg = R.OracleConfig(batch_size=6, max_caption_length=20)
    w = R.init_weights(cfg, seed=3, random_bias=False)
    feats = _fixture()[0]
    out = R.train_forward(cfg, w, feats[:6], sent[:6], masks[:6], np.float64)
    assert abs(out["attentions"].mean() * 196 - masks[:6].astype(np.float64).sum(1).mean()) < 1e-9


def _oracle_step(ctx, sent, masks, seed):
    cfg = R.OracleConfig(batch_size=ctx.shape[0])
    w = R.init_weights(cfg, seed=seed, random_bias=False)          # reference init: U(-0.08, 0.08), zero biases
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        losses, g = TR.loss_and_grads(cfg, w, ctx, sent, masks, seed=seed + 10, reg_in_grad=True)   # dropout ON
    norm = float(np.sqrt(sum(float((x ** 2).sum()) for x in g.values())))
    return losses, norm


def test_backward_oracle_brackets_the_recorded_cross_entropy_and_gradient_norm():
    """B=20, default graph, dropout on, real captions.  The feature magnitude of the recorded batch is unknown (the
    step was taken on whatever the CNN produced then), and both numbers grow monotonically with it: over a range of
    feature scales the oracle's cross entropy passes through the recorded 8.7459 and its gradient norm through the
    recorded 2.5466, and the coverage and regulariser terms match throughout.  A wrong backward pass (a missing
    time step, a dropout scale, the 1/sum(masks) normaliser) moves the norm by factors, not by the +-25 % bracket."""
    _, sent_all, masks_all, _ = _fixture()
    cfg = R.OracleConfig(batch_size=20)
    rng = np.random.RandomState(7)
    pick = rng.permutation(40)[:20]
    sent, masks = sent_all[pick], masks_all[pick]
    base = R.synth_contexts(cfg, 20, seed=7)
    lo_l, lo_n = _oracle_step(base * np.float32(0.2), sent, masks, 7)
    hi_l, hi_n = _oracle_step(base * np.float32(0.6), sent, masks, 7)
    assert lo_l["cross_entropy_loss"] < REC_CE < hi_l["cross_entropy_loss"]
    assert lo_n < REC_NORM < hi_n
    assert 0.75 * REC_NORM < 0.5 * (lo_n + hi_n) < 1.25 * REC_NORM
    for l in (lo_l, hi_l):
        assert abs(l["attention_loss"] - REC_ATT) < 0.03 * REC_ATT
        assert abs(l["reg_loss"] - S["metrics/reg_loss"]) < 0.01 * S["metrics/reg_loss"]
        assert l["accuracy"] == S["metrics/accuracy"] == 0.0
        assert abs(l["total_loss"] - (l["cross_entropy_loss"] + l["attention_loss"] + l["reg_loss"])) < 1e-9
    # clip_by_global_norm(5.0) left the recorded gradient untouched (clipped == unclipped), as the oracle's clip does
    assert S["optimizer/OptimizeLoss/global_norm/clipped_gradient_norm"] == REC_NORM < 5.0


def test_real_conv5_3_features_saturate_but_keep_the_loss_identities():
    """With the REAL conv5_3 features of the images the reference ships (94 % zeros, values up to ~290) the random
    init decoder is far more excited than in the recorded step (cross entropy ~10.3, gradient norm ~11, clipped by
    5.0): recorded here so that the difference to the recorded step is on file, not hidden."""
    feats, sent_all, masks_all, stats = _fixture()
    assert 0.9 < stats["zero_fraction"] < 0.97 and stats["max"] > 100
    idx = list(range(14)) + [0, 3, 5, 7, 9, 11]
    l, n = _oracle_step(feats[idx], sent_all[:20], masks_all[:20], 1)
    assert 9.5 < l["cross_entropy_loss"] < 11.5 and 5.0 < n < 20.0
    assert abs(l["attention_loss"] - REC_ATT) < 0.03 * REC_ATT
