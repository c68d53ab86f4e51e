"""GPU: training-side pool evaluation (BASELINE config 5) against the oracle, which is pinned to the
reference's own StageClassifier::Evaluate (tests/test_oracle_vs_ref.py::test_pool_eval)."""
import numpy as np
import pytest

from oracle import oracle as O
from surfcascade_b200 import synth

pytestmark = pytest.mark.gpu


def _descriptors(N, n_pos, sel):
    pool = O.pool_patches(40)
    X = np.zeros((N, len(sel), 32), np.float32)
    for n in range(N):
        img = synth.positive(n) if n < n_pos else np.ascontiguousarray(synth.negative_frame(n)[7:47, 11:51])
        X[n] = O.features(O.integral(img), pool[sel])[0]
    return X


@pytest.mark.parametrize("N,n_pos,P,T", [(64, 32, 32, 0), (203, 97, 45, 2), (500, 250, 608, 3)])
def test_pool_eval_matches_oracle(gpu_handle, N, n_pos, P, T):
    rng = np.random.default_rng(N)
    sel = np.arange(608) if P == 608 else np.linspace(0, 607, P).astype(int)
    X = _descriptors(min(N, 120), min(n_pos, 60), sel)
    if N > len(X):  # enlarge with jittered copies: the kernel sees distinct rows, the oracle the same ones
        reps = -(-N // len(X))
        pos, neg = X[:min(n_pos, 60)], X[min(n_pos, 60):]
        Xp = np.concatenate([pos] * reps)[:n_pos]
        Xn = np.concatenate([neg] * reps)[:N - n_pos]
        X = np.concatenate([Xp, Xn]) + rng.normal(0, 0.01, size=(N, P, 32)).astype(np.float32)
    W = rng.normal(0, 1.2, size=(P, 33)).astype(np.float32)
    b = np.ones(P)
    prior = None
    if T:
        prior = np.zeros(N, np.float32)
        for t in range(T):
            wt = rng.normal(0, 1, 33).astype(np.float32)
            for n in range(N):
                prior[n] = np.float32(prior[n] + np.float32(O.weak(wt, 1.0, X[n, (7 * t) % P])))
    labels = np.zeros(N, np.uint8)
    labels[:n_pos] = 1
    want = O.pool_eval(X, n_pos, W, b, prior, T)
    got = gpu_handle.pool_eval(X, labels, W, b, prior, T)
    # the histogram is exact integer work; a score landing on the other side of a threshold because of the double exp's
    # last ulp would move the AUC by 1/N: allow none in general, tolerate 1e-5 relative per north_star
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-7)
    assert (got.view(np.uint32) == want.view(np.uint32)).mean() > 0.99


def test_pool_eval_sharded_histograms_add_up(gpu_handle):
    """Sample shards (the multi-GPU partition of config 5) accumulate into one histogram: same AUC as one pass."""
    import torch
    rng = np.random.default_rng(5)
    N, n_pos, P = 96, 40, 64
    X = rng.normal(0, 0.2, size=(N, P, 32)).astype(np.float32)
    labels = np.zeros(N, np.uint8); labels[:n_pos] = 1
    perm = rng.permutation(N)  # shards need not keep positives first
    W = rng.normal(0, 1.0, size=(P, 33)).astype(np.float32); b = np.ones(P)
    whole = gpu_handle.pool_eval(X, labels, W, b)
    hist = torch.zeros(P * 2 * 21, dtype=torch.int32, device="cuda:0")
    for shard in np.array_split(perm, 3):
        dX = torch.from_numpy(np.ascontiguousarray(X[shard])).cuda()
        dl = torch.from_numpy(np.ascontiguousarray(labels[shard])).cuda()
        torch.cuda.synchronize()
        gpu_handle.pool_hist_device(dX.data_ptr(), len(shard), P, dl.data_ptr(), W, b, None, 0, hist.data_ptr())
        gpu_handle.sync()
    assert int(hist.sum().item()) == N * P
    got = gpu_handle.pool_auc_device(hist.data_ptr(), P, n_pos, N - n_pos)
    assert np.array_equal(got.view(np.uint32), whole.view(np.uint32))


def test_extract_pool_features_bit_exact(gpu_handle):
    """Next row N3: descriptors of all 608 pool patches of a batch of 40x40 samples (ExtractNextImageFeatures) == oracle."""
    from oracle import oracle as O
    from surfcascade_b200 import synth
    rng = np.random.default_rng(5)
    imgs = np.stack([synth.positive(s) for s in range(6)] + [rng.integers(0, 256, (40, 40), dtype=np.uint8) for _ in range(3)]
                    + [np.zeros((40, 40), np.uint8), np.full((40, 40), 255, np.uint8)])
    X = gpu_handle.extract_pool_features(imgs)
    pool = O.pool_patches(40)
    assert X.shape == (len(imgs), len(pool), 32)
    for i, img in enumerate(imgs):
        want, _ = O.features(O.integral(img), pool)
        assert np.array_equal(X[i].view(np.uint32), want.view(np.uint32)), i
