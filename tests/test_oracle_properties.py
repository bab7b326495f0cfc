"""Property tests of the CPU oracle (hypothesis): the size-independent invariants the GPU parity tests also rely on
(SURVEY.md §4).  CPU only; the oracle is the checker, so its own consistency is worth pinning independently of the
reference fixtures."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import ssl_oracle as O


def _rand(seed, *shape):
    return np.random.default_rng(seed).standard_normal(shape)


@settings(max_examples=25, deadline=None)
@given(n=st.integers(2, 24), d=st.integers(2, 16), seed=st.integers(0, 10_000), tau=st.sampled_from([0.07, 0.5, 1.0]))
def test_ntxent_permutation_invariance_and_orthogonality(n, d, seed, tau):
    zi, zj = _rand(seed, n, d), _rand(seed + 1, n, d)
    loss, dzi, dzj = O.ntxent(zi, zj, True, tau)
    perm = np.random.default_rng(seed + 2).permutation(n)
    loss_p, dzi_p, _ = O.ntxent(zi[perm], zj[perm], True, tau)
    assert abs(loss - loss_p) < 1e-10 * max(1.0, abs(loss))
    assert np.allclose(dzi[perm], dzi_p, rtol=1e-9, atol=1e-12)
    # normalised inputs: the gradient is orthogonal to every row (scale invariance)
    assert np.abs((dzi * zi).sum(1)).max() < 1e-10 and np.abs((dzj * zj).sum(1)).max() < 1e-10
    # swapping the two views leaves the loss unchanged and swaps the gradients
    loss_s, dzj_s, dzi_s = O.ntxent(zj, zi, True, tau)
    assert abs(loss - loss_s) < 1e-10 * max(1.0, abs(loss)) and np.allclose(dzi, dzi_s, rtol=1e-9, atol=1e-12)


@settings(max_examples=25, deadline=None)
@given(b=st.integers(2, 40), k=st.integers(2, 24), iters=st.integers(1, 5), seed=st.integers(0, 10_000))
def test_sinkhorn_marginals(b, k, iters, seed):
    z = _rand(seed, b, 8)
    c = _rand(seed + 1, k, 8)
    z /= np.linalg.norm(z, axis=1, keepdims=True)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    q = O.sinkhorn(z @ c.T, 0.05, iters)
    assert q.shape == (b, k) and (q >= 0).all()
    np.testing.assert_allclose(q.sum(1), 1.0, rtol=1e-9)          # rows (samples) sum to one
    assert abs(q.sum() - b) < 1e-9 * b


@settings(max_examples=25, deadline=None)
@given(size=st.integers(1, 12), batches=st.lists(st.integers(0, 30), min_size=1, max_size=5), seed=st.integers(0, 10_000),
       normalize=st.booleans())
def test_ring_matches_sequential_loop(size, batches, seed, normalize):
    """models/moco.py:31-36 / models/swav.py:70-75: row by row, pointer wraps, last writer wins when n > size."""
    d = 4
    bank, ptr = np.zeros((size, d), np.float32), 0
    ref, rptr = np.zeros((size, d), np.float32), 0
    for j, n in enumerate(batches):
        x = _rand(seed + j, n, d).astype(np.float32)
        bank, ptr = O.ring_enqueue(bank, ptr, x, normalize)
        for r in range(n):
            row = x[r]
            if normalize:
                row = (row / max(float(np.sqrt((row.astype(np.float64) ** 2).sum())), 1e-12)).astype(np.float32)
            ref[rptr] = row
            rptr = (rptr + 1) % size
        assert ptr == rptr
        np.testing.assert_allclose(bank, ref, rtol=3e-7, atol=0)


@settings(max_examples=20, deadline=None)
@given(n=st.integers(3, 20), d=st.integers(2, 12), seed=st.integers(0, 10_000), lm=st.sampled_from([0.005, 0.05]))
def test_barlow_gradient_matches_finite_differences(n, d, seed, lm):
    zi, zj = _rand(seed, n, d), _rand(seed + 1, n, d)
    loss, dzi, dzj = O.barlow(zi, zj, False, lm)
    rng = np.random.default_rng(seed + 2)
    for _ in range(3):
        r, c = rng.integers(n), rng.integers(d)
        e = np.zeros_like(zi)
        e[r, c] = 1e-6
        num = (O.barlow(zi + e, zj, False, lm)[0] - O.barlow(zi - e, zj, False, lm)[0]) / 2e-6
        assert abs(num - dzi[r, c]) <= 1e-5 * max(1.0, abs(dzi[r, c]), abs(num))
    # standardisation makes the loss invariant to per-column affine maps of the inputs
    a, b = rng.uniform(0.5, 2.0, d), rng.standard_normal(d)
    assert abs(O.barlow(zi * a + b, zj, False, lm)[0] - loss) < 1e-8 * max(1.0, abs(loss))


@settings(max_examples=20, deadline=None)
@given(bs=st.integers(1, 6), nv=st.integers(1, 5), k=st.integers(2, 12), seed=st.integers(0, 10_000))
def test_dino_gradient_rows_sum_to_zero_and_fd(bs, nv, k, seed):
    t, s, c = _rand(seed, bs, 2, k), _rand(seed + 1, bs, nv, k), 0.1 * _rand(seed + 2, k)
    loss, ds = O.dino(t, s, 0.1, 0.04, c)
    assert np.abs(ds.sum(-1)).max() < 1e-10          # softmax-CE gradient: every row sums to zero
    e = np.zeros_like(s)
    e[0, 0, 0] = 1e-6
    num = (O.dino(t, s + e, 0.1, 0.04, c)[0] - O.dino(t, s - e, 0.1, 0.04, c)[0]) / 2e-6
    assert abs(num - ds[0, 0, 0]) <= 1e-5 * max(1.0, abs(num))


@settings(max_examples=20, deadline=None)
@given(n=st.integers(4, 40), d=st.integers(2, 12), seed=st.integers(0, 10_000), lm=st.sampled_from([0.005, 0.05]))
def test_barlow_closed_form_standardisation_backward(n, d, seed, lm):
    """The identity the fused Barlow backward relies on (csrc/barlow.cu, `barlow_fwd_finish_kernel` / `EPI_BARLOW_BWD`):
    with x~ standardised and dT = dL/dx~, mean_n(dT) = 0 and sum_n(dT x~) = row / column sums of dC .* C, so
    dx = (dT - x~ * rowsum(dC .* C) / (n - 1)) / std needs no pass over dT.  Checked against the oracle's gradient, which
    follows the reference's autograd form (utils/losses.py:136-142)."""
    zi = _rand(seed, n, d) * 1.3 + 0.2
    zj = 0.7 * zi + 0.3 * _rand(seed + 1, n, d)
    _, dzi, dzj = O.barlow(zi, zj, False, lm)
    sdi, sdj = zi.std(0, ddof=1), zj.std(0, ddof=1)
    ti, tj = (zi - zi.mean(0)) / sdi, (zj - zj.mean(0)) / sdj
    c = ti.T @ tj / n
    w = np.full((d, d), lm)
    np.fill_diagonal(w, 1.0)
    dc = 2.0 * (c - np.eye(d)) * w
    dti, dtj = tj @ dc.T / n, ti @ dc / n
    assert np.abs(dti.mean(0)).max() < 1e-12 and np.abs(dtj.mean(0)).max() < 1e-12
    m2_i, m2_j = (dc * c).sum(1) / (n - 1), (dc * c).sum(0) / (n - 1)
    np.testing.assert_allclose((dti * ti).sum(0) / (n - 1), m2_i, rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose((dtj * tj).sum(0) / (n - 1), m2_j, rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose((dti - ti * m2_i) / sdi, dzi, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose((dtj - tj * m2_j) / sdj, dzj, rtol=1e-8, atol=1e-12)


@settings(max_examples=20, deadline=None)
@given(b=st.integers(2, 30), k=st.integers(2, 20), iters=st.integers(1, 5), seed=st.integers(0, 10_000))
def test_sinkhorn_codes_from_scaling_vectors(b, k, iters, seed):
    """The identity the fused SwAV cross-entropy relies on (csrc/swav.cu `swav_ce_sk4_kernel`, csrc/sinkhorn.cu): in
    scaling-vector form E = exp((s - max)/eps), alpha_k <- (1/K) / sum_b E_bk beta_b, beta_b <- (1/B) / sum_k alpha_k E_bk,
    the reference's codes (utils/losses.py:213-224) are alpha_k E_bk / sum_k alpha_k E_bk with the LAST alpha - no final
    pass over a stored Q is needed."""
    z = _rand(seed, b, 8)
    c = _rand(seed + 1, k, 8)
    z /= np.linalg.norm(z, axis=1, keepdims=True)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    s = z @ c.T
    eps = 0.05
    e = np.exp((s - s.max()) / eps)
    beta = np.ones(b)
    for _ in range(iters):
        alpha = (1.0 / k) / (e * beta[:, None]).sum(0)
        beta = (1.0 / b) / (e * alpha[None, :]).sum(1)
    codes = e * alpha[None, :]
    codes /= codes.sum(1, keepdims=True)
    np.testing.assert_allclose(codes, O.sinkhorn(s, eps, iters), rtol=1e-9, atol=1e-15)
