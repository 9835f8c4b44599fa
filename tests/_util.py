"""Shared helpers of the GPU parity tests: build a sat_b200.CaptionGenerator and the matching oracle."""
import numpy as np

from oracle import ref_step as R

# parity bar of the north star: 1e-3 relative to the fp32 reference, stated per tensor as
#   max|x - ref| <= TOL * max|ref|   (many logits are near 0, so a pure
#   element-wise rtol would be meaningless)
TOL = 1e-3


def rel_err(got, ref):
    ref = np.asarray(ref, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


def assert_close(got, ref, name, tol=TOL):
    e = rel_err(got, ref)
    assert e <= tol, "%s: max-norm relative error %.3e > %.1e" % (name, e, tol)
    return e


def make_pair(batch, beam=1, seed=1234, max_batch=None, **dims):
    """(oracle cfg, weights, model) with identical shapes; dims use OracleConfig field names."""
    import sat_b200
    ocfg = R.OracleConfig(batch_size=batch, beam_size=beam, **dims)
    w = R.init_weights(ocfg, seed=seed)
    cfg = sat_b200.Config(
        batch_size=batch, beam_size=beam, num_ctx=ocfg.num_ctx, dim_ctx=ocfg.dim_ctx,
        dim_embedding=ocfg.dim_embedding, num_lstm_units=ocfg.num_lstm_units,
        num_initalize_layers=ocfg.num_initalize_layers, dim_initalize_layer=ocfg.dim_initalize_layer,
        num_attend_layers=ocfg.num_attend_layers, dim_attend_layer=ocfg.dim_attend_layer,
        num_decode_layers=ocfg.num_decode_layers, dim_decode_layer=ocfg.dim_decode_layer,
        vocabulary_size=ocfg.vocabulary_size, max_caption_length=ocfg.max_caption_length)
    model = sat_b200.CaptionGenerator(cfg, max_batch=max_batch)
    assert model.set_weights(w) == 0
    return ocfg, w, model


SMALL = dict(num_ctx=49, dim_ctx=64, dim_embedding=32, num_lstm_units=64, dim_initalize_layer=32,
             dim_attend_layer=32, dim_decode_layer=64, vocabulary_size=300, max_caption_length=6)
