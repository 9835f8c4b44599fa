"""Host-side logic that needs no GPU: config mirror, variable naming, batch sharding, and the
world_size-2 gloo path of the multi-GPU plumbing."""
import os
import socket

import numpy as np
import pytest

import sat_b200
from oracle import ref_step as R
from sat_b200 import parallel


def test_config_field_names_follow_the_reference():
    c = sat_b200.Config()
    for f in ("max_caption_length", "dim_embedding", "num_lstm_units", "num_initalize_layers",
              "dim_initalize_layer", "num_attend_layers", "dim_attend_layer", "num_decode_layers",
              "dim_decode_layer", "vocabulary_size", "batch_size", "beam_size", "fc_drop_rate", "lstm_drop_rate"):
        assert hasattr(c, f)
    assert (c.num_ctx, c.dim_ctx) == (196, 512)
    r = sat_b200.Config(cnn="resnet50")
    assert (r.num_ctx, r.dim_ctx) == (49, 2048)          # model.py:103-108
    with pytest.raises(AttributeError):
        sat_b200.Config(nonsense=1)


@pytest.mark.parametrize("layers", [1, 2])
def test_facade_variable_table_equals_oracle(layers):
    c = sat_b200.Config(num_attend_layers=layers, num_decode_layers=layers, num_initalize_layers=layers)
    o = R.OracleConfig(num_attend_layers=layers, num_decode_layers=layers, num_initalize_layers=layers)
    assert sat_b200.weight_shapes(c) == R.weight_shapes(o)


def test_shard_range_partitions_exactly():
    for n in (1, 7, 64, 512):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def test_dropout_generator_matches_the_numpy_copy(built_lib):
    """the C generator used by the training kernels == oracle/train_ref.uniform24 (runs on the host)"""
    from oracle import train_ref as TR
    lib = sat_b200.load_library()
    for seed, stream in ((1234, 5), (77, 16 * 19 + 7), (2 ** 40 + 3, 0xFFFF2)):
        ref = TR.uniform24(seed, stream, 64)
        got = np.array([lib.sat_train_rng_uniform(seed, stream, i) for i in range(64)], np.float32)
        assert np.array_equal(ref, got)
    big = TR.uniform24(9, 1, 1 << 20)
    assert lib.sat_train_rng_uniform(9, 1, (1 << 20) - 1) == big[-1]


def test_dropout_keep_threshold_is_the_float_formula(tmp_path):
    """The training kernels decide "kept" with one integer compare (sat::DropGen); the threshold must reproduce
    floor(keep + u) in float32 for every 24-bit u.  Exhaustive, on the host (nvcc compiles the header for the CPU)."""
    import shutil, subprocess
    from oracle import train_ref as TR
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "drop_threshold_check.cu")
    exe = str(tmp_path / "drop_threshold_check")
    subprocess.run([nvcc, "-O2", "-o", exe, src], check=True, capture_output=True)
    keeps = ["0.5", "0.7", "0.3", "0.9", "1.0", "0.1", "0.999", "0.33333334"]
    out = subprocess.run([exe] + keeps, check=True, capture_output=True, text=True).stdout.splitlines()
    assert len(out) == len(keeps) + 1 and all(l.endswith("mismatches 0") for l in out[:-1]), out
    u = np.array([float(x) for x in out[-1].split()[1:]], np.float32)   # printed with 9 significant digits: exact in float32
    assert np.array_equal(u[:2], TR.uniform24(1234, 5, 2))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    parallel.init_process_group("gloo")
    lo, hi = parallel.shard_range(n, rank, world)
    # each rank "decodes" its shard of images: token = image id * 10 + t (stands in for the GPU step)
    local = torch.tensor([[i * 10 + t for t in range(4)] for i in range(lo, hi)], dtype=torch.int32)
    full = parallel.gather_shards(local, n)
    tmax = parallel.max_over_ranks(1.0 + rank)
    q.put((rank, full.numpy().tolist(), tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_and_max():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n = 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [[i * 10 + t for t in range(4)] for i in range(n)]
    for rank, full, tmax in outs:
        assert full == expect
        assert tmax == 2.0


def _dp_worker(rank, world, port, q):
    """Two data-parallel training steps over gloo: the oracle stands in for each rank's forward + backward pass, the
    host protocol (parallel.StepCollective: one collective per step, next batch's mask sum riding with the gradients)
    is the product code under test."""
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import warnings
    import torch
    import torch.distributed as dist
    from oracle import ref_step as R, train_ref as TR
    parallel.init_process_group("gloo")
    cfg = R.OracleConfig(batch_size=4, num_ctx=6, dim_ctx=8, dim_embedding=6, num_lstm_units=8, dim_initalize_layer=6,
                         dim_attend_layer=6, dim_decode_layer=8, vocabulary_size=20, max_caption_length=4)
    w = R.init_weights(cfg, seed=2)
    names = sorted(w)
    n_grad = sum(w[k].size for k in names)
    dp = parallel.StepCollective(torch.zeros(n_grad + 8, dtype=torch.float32), n_grad)
    rng = np.random.RandomState(5)
    batches = []
    for it in range(3):
        ctx = R.synth_contexts(cfg, 4, seed=10 + it)
        sent = rng.randint(1, 20, (4, 4)).astype(np.int32)
        lens = rng.randint(1, 5, 4)
        batches.append((ctx, sent, (np.arange(4)[None, :] < lens[:, None]).astype(np.float32)))
    lo, hi = parallel.shard_range(4, rank, world)
    out = []
    for it in range(2):
        ctx, sent, masks = batches[it]
        mk = torch.tensor(masks[lo:hi])
        nxt = torch.tensor(batches[it + 1][2][lo:hi])
        gsum = float(dp.global_mask_sum(mk).item())
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            l, g = TR.loss_and_grads(cfg, w, ctx[lo:hi], sent[lo:hi], masks[lo:hi], None, global_mask_sum=gsum,
                                     global_batch=4, reg_in_grad=False)
        dp.flat[:n_grad].copy_(torch.tensor(np.concatenate([g[k].ravel() for k in names]), dtype=torch.float32))
        shard = torch.tensor([l["cross_entropy_loss"], l["accuracy"], l["attention_loss"], 0.0], dtype=torch.float32)
        glob = dp.reduce(shard, nxt, announced=True)
        out.append((gsum, dp.flat[:n_grad].numpy().copy(), glob.numpy().copy(), float(dp.mask_sum.item())))
    q.put((rank, dp.collectives, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_training_step_protocol():
    import warnings
    import torch.multiprocessing as mp
    from oracle import ref_step as R, train_ref as TR
    ctx_ = mp.get_context("spawn")
    q = ctx_.Queue()
    port = _free_port()
    procs = [ctx_.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the same two batches on ONE process
    cfg = R.OracleConfig(batch_size=4, num_ctx=6, dim_ctx=8, dim_embedding=6, num_lstm_units=8, dim_initalize_layer=6,
                         dim_attend_layer=6, dim_decode_layer=8, vocabulary_size=20, max_caption_length=4)
    w = R.init_weights(cfg, seed=2)
    names = sorted(w)
    rng = np.random.RandomState(5)
    for it in range(2):
        ctx = R.synth_contexts(cfg, 4, seed=10 + it)
        sent = rng.randint(1, 20, (4, 4)).astype(np.int32)
        lens = rng.randint(1, 5, 4)
        masks = (np.arange(4)[None, :] < lens[:, None]).astype(np.float32)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            l, g = TR.loss_and_grads(cfg, w, ctx, sent, masks, None, reg_in_grad=False)
        ref = np.concatenate([g[k].ravel() for k in names])
        for rank, ncoll, steps in outs:
            gsum, flat, glob, nxt_sum = steps[it]
            assert gsum == float(masks.sum())                       # the normaliser every rank used is the global one
            assert np.abs(flat - ref).max() < 1e-5 * np.abs(ref).max()
            assert abs(glob[0] - l["cross_entropy_loss"]) < 1e-5 * l["cross_entropy_loss"]
            assert abs(glob[2] - l["attention_loss"]) < 1e-5 * l["attention_loss"]
            assert abs(glob[1] - l["accuracy"]) < 1e-6
    for rank, ncoll, steps in outs:
        assert ncoll == 3            # mask sum of the first batch + ONE collective per step: the second step needed no extra one
        assert steps[0][3] == steps[1][0]


def test_golden_fixtures_match_current_oracle():
    """tests/golden/step_*.npz were produced by the fp64 oracle; guard against silent drift."""
    here = os.path.join(os.path.dirname(__file__), "golden")
    for tag, layers in (("2layer", 2), ("1layer", 1)):
        z = np.load(os.path.join(here, "step_%s.npz" % tag))
        cfg = R.OracleConfig(num_ctx=49, dim_ctx=64, dim_embedding=32, num_lstm_units=64, dim_initalize_layer=32,
                             dim_attend_layer=32, dim_decode_layer=64, vocabulary_size=300, batch_size=3,
                             max_caption_length=6, num_attend_layers=layers, num_decode_layers=layers,
                             num_initalize_layers=layers)
        w = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
        r = R.decode_step(cfg, w, z["ctx"], z["last_word"], z["last_memory"], z["last_output"], np.float64)
        for k in ("memory", "output", "probs", "logits", "alpha", "context"):
            np.testing.assert_allclose(r[k], z["step_" + k], rtol=1e-12, atol=1e-14)
        toks, _ = R.decode_loop(cfg, w, z["ctx"], 6, z["forced"], np.float64)
        assert (toks == z["tokens"]).all()
