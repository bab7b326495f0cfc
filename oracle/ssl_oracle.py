"""CPU oracle for the self-supervised loss hot path (TEST INFRASTRUCTURE ONLY).

This module is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The shipped path (``ssv_b200``) never does and
raises when its CUDA library is missing.

It is an independent numpy-fp64 closed-form restatement (loss AND analytic
gradients, no autograd) of the reference's PyTorch losses.  Each function
cites the reference file:line it follows (paths relative to the reference
repo root).  The arithmetic the reference relies on lives in PyTorch ATen
(torch==1.8.1 pinned in requirements.txt:5; semantics used here are stable up
to the torch 2.11 of this image): ``F.normalize`` eps clamp 1e-12,
``torch.std`` unbiased, ``F.cross_entropy`` mean reduction,
``F.kl_div(log_target=True, reduction='sum')`` = sum(exp(t) * (t - x)).

Parity pinning: the reference ships NO tests / golden vectors for this path
(SURVEY.md §4, §8c), so the oracle is pinned against outputs of the reference
itself: ``tests/golden/make_golden.py`` imports ``/root/reference/utils/losses.py``
(and the bank classes of ``models/moco.py`` / ``models/swav.py``), runs it with
torch autograd on seeded inputs and commits inputs + outputs as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
below against those fixtures.
"""
from __future__ import annotations

import numpy as np

F64 = np.float64
NORM_EPS = 1e-12  # F.normalize default eps


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
def _f64(x):
    return np.asarray(x, dtype=F64)


def l2_normalize(x):
    """F.normalize(x, p=2, dim=-1): x / max(||x||, 1e-12).  Returns (xhat, denom)."""
    x = _f64(x)
    nrm = np.sqrt((x * x).sum(-1, keepdims=True))
    den = np.maximum(nrm, NORM_EPS)
    return x / den, den


def l2_normalize_bwd(dxhat, xhat, den):
    """Backward of l2_normalize.  (For ||x|| < eps the clamp branch is linear;
    the projection term then vanishes with xhat ~ 0, matching autograd to O(eps).)"""
    return (dxhat - (dxhat * xhat).sum(-1, keepdims=True) * xhat) / den


def _lse(x, axis=-1):
    m = x.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(x - m).sum(axis=axis, keepdims=True))).squeeze(axis)


# --------------------------------------------------------------------------
# a1: SimCLR NT-Xent   (utils/losses.py:15-46)
# --------------------------------------------------------------------------
def ntxent(zi, zj, normalize=False, temperature=1.0):
    """Returns (loss, dzi, dzj).

    utils/losses.py:20-25 optional normalise; :27-30 four NxN logits blocks == the
    2Nx2N matrix S = Z Z^T / tau with Z = [zi; zj]; :32-44 positive is S[a, a+-N],
    negatives are every other off-diagonal entry; :45 mean cross-entropy over 2N rows.
    """
    zi, zj = _f64(zi), _f64(zj)
    n = zi.shape[0]
    z = np.concatenate([zi, zj], 0)
    m = 2 * n
    if normalize:
        zh, den = l2_normalize(z)
    else:
        zh, den = z, None
    s = zh @ zh.T / temperature
    idx = np.arange(m)
    partner = (idx + n) % m
    s_masked = s.copy()
    s_masked[idx, idx] = -np.inf
    lse = _lse(s_masked, 1)
    pos = s[idx, partner]
    loss = float((lse - pos).mean())
    # backward: G = (P - Y)/M with P the masked softmax; S is symmetric in Z so
    # dZhat = (G + G^T) Zhat / tau
    p = np.exp(s_masked - lse[:, None])
    g = p
    g[idx, partner] -= 1.0
    g /= m
    dzh = (g + g.T) @ zh / temperature
    dz = l2_normalize_bwd(dzh, zh, den) if normalize else dzh
    return loss, dz[:n], dz[n:]


# --------------------------------------------------------------------------
# a2: MoCo InfoNCE   (utils/losses.py:56-72)
# --------------------------------------------------------------------------
def moco(query, keys, memory, normalize=True, temperature=1.0):
    """Returns (loss, dquery, dkeys).  Queue rows are used as stored
    (utils/losses.py:69: no re-normalisation, no gradient)."""
    q, k, mem = _f64(query), _f64(keys), _f64(memory)
    n = q.shape[0]
    if normalize:
        qh, qd = l2_normalize(q)
        kh, kd = l2_normalize(k)
    else:
        qh, kh, qd, kd = q, k, None, None
    pos = (qh * kh).sum(1) / temperature           # :68 diagonal of q k^T
    neg = qh @ mem.T / temperature                 # :69
    logits = np.concatenate([pos[:, None], neg], 1)  # :70
    lse = _lse(logits, 1)
    loss = float((lse - pos).mean())               # :71 CE with label 0
    p = np.exp(logits - lse[:, None])
    p0 = p[:, 0]
    dqh = ((p0 - 1.0)[:, None] * kh + p[:, 1:] @ mem) / (n * temperature)
    dkh = (p0 - 1.0)[:, None] * qh / (n * temperature)
    if normalize:
        return loss, l2_normalize_bwd(dqh, qh, qd), l2_normalize_bwd(dkh, kh, kd)
    return loss, dqh, dkh


# --------------------------------------------------------------------------
# a3 / a7: ring buffers   (models/moco.py:23-39, models/swav.py:57-79)
# --------------------------------------------------------------------------
def ring_enqueue(bank, ptr, batch, normalize):
    """Row-by-row ring write, exactly the loop of models/moco.py:31-36
    (normalize=True, fp32 x / max(||x||,1e-12)) and models/swav.py:70-75
    (normalize=False).  Operates in float32 like the reference; returns
    (new_bank, new_ptr).  ptr / which-row-overwritten semantics are bit-exact:
    wraps mid-batch, and when len(batch) > size the last writer wins."""
    bank = np.array(bank, dtype=np.float32, copy=True)
    size = bank.shape[0]
    batch = np.asarray(batch, dtype=np.float32)
    for row in batch:
        if normalize:
            nrm = np.sqrt(np.sum(row.astype(np.float32) ** 2, dtype=np.float32))
            row = row / np.maximum(nrm, np.float32(NORM_EPS))
        bank[ptr] = row
        ptr += 1
        if ptr >= size:
            ptr = 0
    return bank, ptr


# --------------------------------------------------------------------------
# a4: Barlow Twins   (utils/losses.py:127-142)
# --------------------------------------------------------------------------
def barlow(zi, zj, normalize=True, lmbda=0.005):
    """Returns (loss, dzi, dzj).  :136-137 standardise with UNBIASED std;
    :138 C = Xi^T Xj / N; :139-142 sum((C-I)^2 * (lambda off-diag, 1 on diag))."""
    xi, xj = _f64(zi), _f64(zj)
    n, d = xi.shape
    if normalize:
        xi_h, di = l2_normalize(xi)
        xj_h, dj = l2_normalize(xj)
    else:
        xi_h, xj_h, di, dj = xi, xj, None, None

    def standardize(x):
        mu = x.mean(0)
        sd = x.std(0, ddof=1)
        return (x - mu) / sd, sd

    ti, sdi = standardize(xi_h)
    tj, sdj = standardize(xj_h)
    c = ti.T @ tj / n
    eye = np.eye(d)
    w = np.full((d, d), lmbda)
    np.fill_diagonal(w, 1.0)
    loss = float((((c - eye) ** 2) * w).sum())
    dc = 2.0 * (c - eye) * w
    dti = tj @ dc.T / n
    dtj = ti @ dc / n

    def standardize_bwd(dt, t, sd):
        return (dt - dt.mean(0) - t * (dt * t).sum(0) / (n - 1)) / sd

    dxi = standardize_bwd(dti, ti, sdi)
    dxj = standardize_bwd(dtj, tj, sdj)
    if normalize:
        dxi = l2_normalize_bwd(dxi, xi_h, di)
        dxj = l2_normalize_bwd(dxj, xj_h, dj)
    return loss, dxi, dxj


# --------------------------------------------------------------------------
# a8: BYOL nn.MSELoss   (models/byol.py:89,129-130)
# --------------------------------------------------------------------------
def mse(inp, target):
    """nn.MSELoss() mean reduction.  Returns (loss, dinp, dtarget)."""
    o, t = _f64(inp), _f64(target)
    diff = o - t
    loss = float((diff * diff).mean())
    g = 2.0 * diff / diff.size
    return loss, g, -g


# --------------------------------------------------------------------------
# a9: SimSiam   (utils/losses.py:150-151)
# --------------------------------------------------------------------------
def simsiam(online, target):
    """-(o*t).sum(1).mean().  Returns (loss, donline, dtarget)."""
    o, t = _f64(online), _f64(target)
    n = o.shape[0]
    loss = float(-(o * t).sum(1).mean())
    return loss, -t / n, -o / n


# --------------------------------------------------------------------------
# a10: ReLIC   (utils/losses.py:162-201)
# --------------------------------------------------------------------------
def relic(zi, zj, zo, normalize=True, temperature=1.0, alpha=0.5):
    """Returns (loss, dzi, dzj, dzo).  Contrastive part == ntxent (:163-194).
    KL quirk reproduced (:196-200): a_n = zi_n.zo_n/tau, b_n = zj_n.zo_n/tau,
    p = softmax_n(a) (probabilities passed as kl_div *input*),
    lq = log_softmax_n(b) (log_target) -> KL = sum_n exp(lq_n) * (lq_n - p_n)."""
    zi, zj, zo = _f64(zi), _f64(zj), _f64(zo)
    n = zi.shape[0]
    closs, dzi_c, dzj_c = ntxent(zi, zj, normalize, temperature)
    if normalize:
        ih, idn = l2_normalize(zi)
        jh, jdn = l2_normalize(zj)
        oh, odn = l2_normalize(zo)
    else:
        ih, jh, oh = zi, zj, zo
    a = (ih * oh).sum(1) / temperature
    b = (jh * oh).sum(1) / temperature
    p = np.exp(a - _lse(a, 0))
    lq = b - _lse(b, 0)
    q = np.exp(lq)
    kl = float((q * (lq - p)).sum())
    # dKL/dp_n = -q_n ; through softmax: da = p*(g - sum(p g)) with g = -q
    da = -(p * q - p * (p * q).sum())
    # dKL/dlq_n = q_n (lq_n - p_n) + q_n ; through log_softmax: db = h - q*sum(h)
    h = q * (lq - p) + q
    db = h - q * h.sum()
    dih = alpha * da[:, None] * oh / temperature
    djh = alpha * db[:, None] * oh / temperature
    doh = alpha * (da[:, None] * ih + db[:, None] * jh) / temperature
    if normalize:
        dzi_k = l2_normalize_bwd(dih, ih, idn)
        dzj_k = l2_normalize_bwd(djh, jh, jdn)
        dzo = l2_normalize_bwd(doh, oh, odn)
    else:
        dzi_k, dzj_k, dzo = dih, djh, doh
    return closs + alpha * kl, dzi_c + dzi_k, dzj_c + dzj_k, dzo


# --------------------------------------------------------------------------
# a5: SwAV Sinkhorn-Knopp   (utils/losses.py:213-224)
# --------------------------------------------------------------------------
def sinkhorn(scores, eps=0.05, n_iters=3):
    """Literal restatement: Q = exp(S/eps)^T; Q /= sum; iterate row (prototype)
    then column (sample) normalisation; final column normalisation; return B x K."""
    s = _f64(scores)
    q = np.exp(s / eps).T
    q = q / q.sum()
    k, b = q.shape
    r = np.ones(k) / k
    c = np.ones(b) / b
    for _ in range(n_iters):
        u = q.sum(1)
        q = q * (r / u)[:, None]
        q = q * (c / q.sum(0))[None, :]
    return (q / q.sum(0, keepdims=True)).T


# --------------------------------------------------------------------------
# a6: SwAV loss   (utils/losses.py:226-235)
# --------------------------------------------------------------------------
def swav(z1, z2, prototypes, bank=None, temperature=0.1, eps=0.05, n_iters=3):
    """Returns (loss, dz1, dz2, dprototypes).  dz1/dz2 cover only the live batch
    rows (bank rows get no gradient: models/swav.py:140 passes a constant tensor)."""
    z1, z2, c = _f64(z1), _f64(z2), _f64(prototypes)
    nb = z1.shape[0]
    if bank is not None:
        bank = _f64(bank)
        z1 = np.concatenate([z1, bank], 0)
        z2 = np.concatenate([z2, bank], 0)
    bp = z1.shape[0]
    s1, s2 = z1 @ c.T, z2 @ c.T
    q1, q2 = sinkhorn(s1, eps, n_iters), sinkhorn(s2, eps, n_iters)
    p1 = s1 / temperature - _lse(s1 / temperature, 1)[:, None]
    p2 = s2 / temperature - _lse(s2 / temperature, 1)[:, None]
    loss = float(-0.5 * ((q1 * p2).sum(1) + (q2 * p1).sum(1)).mean())
    # d loss / d s2 = -0.5/B' * (q1 - softmax(s2/T) * rowsum(q1)) / T ; rowsum(q)=1
    ds2 = -0.5 / bp * (q1 - np.exp(p2) * q1.sum(1, keepdims=True)) / temperature
    ds1 = -0.5 / bp * (q2 - np.exp(p1) * q2.sum(1, keepdims=True)) / temperature
    dz1 = ds1 @ c
    dz2 = ds2 @ c
    dc = ds1.T @ z1 + ds2.T @ z2
    return loss, dz1[:nb], dz2[:nb], dc


def prototypes_forward(embedding):
    """models/swav.py:51-54: rows of the embedding table L2-normalised each call."""
    return l2_normalize(embedding)[0]


# --------------------------------------------------------------------------
# (f) next rows: DinoLoss (utils/losses.py:75-89) + centre EMA (models/dino.py:136-141) + parameter EMA
# --------------------------------------------------------------------------
def _log_softmax(x):
    return x - _lse(x, -1)[..., None]


def dino(teacher, student, temp_s, temp_t, center):
    """utils/losses.py:80-89.  teacher [bs,2,K], student [bs,nv,K], center [K] -> (loss, dstudent).
    targets_g = softmax((teacher[:, g] - center)/temp_t) broadcast over the nv student views (:83-86);
    loss = sum_g -mean_{b,v} sum_k targets_g * log_softmax(student/temp_s) (:87-89).  No gradient reaches the teacher
    (models/dino.py:151-152 runs it under no_grad)."""
    t, s, c = _f64(teacher), _f64(student), _f64(center)
    bs, nv, _ = s.shape
    tg = np.exp(_log_softmax((t - c) / temp_t))          # [bs, 2, K]
    tsum = tg.sum(1)[:, None, :]                          # [bs, 1, K]
    logp = _log_softmax(s / temp_s)
    loss = -(tsum * logp).sum() / (bs * nv)
    dstudent = (2.0 * np.exp(logp) - tsum) / (temp_s * bs * nv)
    return loss, dstudent


def dino_center_update(center, teacher_rows, m):
    """models/dino.py:136-141 in fp32 (bit-level restatement: separate roundings of the two products and the sum)."""
    mean = np.asarray(teacher_rows, dtype=np.float32).mean(0, dtype=np.float32)
    if center is None:
        return mean
    return np.float32(m) * np.asarray(center, np.float32) + np.float32(1.0 - m) * mean


def ema_update(target, source, m):
    """models/moco.py:108-111 (byol.py:120-123, relic.py:119-122, dino.py:129-134): t = m*t + (1-m)*s in fp32 with the
    reference's roundings (scalar (1-m) formed in double, then each product and the sum rounded to fp32)."""
    t, s = np.asarray(target, np.float32), np.asarray(source, np.float32)
    return np.float32(m) * t + np.float32(1.0 - m) * s


def pirl(img, patch, mem_pos, mem_neg, normalize=True, temperature=1.0, loss_weight=0.5):
    """utils/losses.py:100-117.  Two cross-entropies (label 0) over [positive | negatives] logits that share the negatives
    mem_pos @ mem_neg^T (:109); head 1 positive = mem_pos . v_patch (:107), head 2 = mem_pos . v_img (:108); the memory
    rows are used as stored.  Returns (loss, d_img, d_patch) (the memory features carry no gradient)."""
    vi, vp, mp, mn = _f64(img), _f64(patch), _f64(mem_pos), _f64(mem_neg)
    n = vi.shape[0]
    if normalize:
        vih, di_ = l2_normalize(vi)
        vph, dp_ = l2_normalize(vp)
    else:
        vih, vph = vi, vp
    neg = mp @ mn.T / temperature
    out = []
    for vh, w in ((vph, loss_weight), (vih, 1.0 - loss_weight)):
        pos = (mp * vh).sum(1) / temperature
        logits = np.concatenate([pos[:, None], neg], 1)
        lse = _lse(logits, -1)
        p0 = np.exp(pos - lse)
        out.append((w * (lse - pos).mean(), w * (p0 - 1.0)[:, None] * mp / (n * temperature)))
    (l1, dph), (l2, dih) = out
    if normalize:
        d_img, d_patch = l2_normalize_bwd(dih, vih, di_), l2_normalize_bwd(dph, vph, dp_)
    else:
        d_img, d_patch = dih, dph
    return l1 + l2, d_img, d_patch


def pirl_bank_update(bank, indices, vectors, m=None):
    """models/pirl.py:32-38 in fp32: m is None -> initialize_vectors (bank[idx] = normalize(v)); else update_vectors
    (bank[idx] = m * bank[idx] + (1 - m) * normalize(v), separate roundings)."""
    bank = np.array(bank, dtype=np.float32, copy=True)
    v = np.asarray(vectors, dtype=np.float32)
    den = np.maximum(np.sqrt((v.astype(np.float32) ** 2).sum(1, dtype=np.float32)), np.float32(1e-12))[:, None]
    vh = (v / den).astype(np.float32)
    idx = np.asarray(indices, dtype=np.int64)
    bank[idx] = vh if m is None else np.float32(m) * bank[idx] + np.float32(1.0 - m) * vh
    return bank


# --------------------------------------------------------------------------
# f4: SeLA self-labelling   (models/sela.py:146-166, per-batch body)
# --------------------------------------------------------------------------
def sela_self_label(logits, alpha, beta, lmbd, num_iters):
    """P = pow(log_softmax(logits, -1), lmbd)^T (:152); num_iters x {alpha = 1/(P beta); beta = 1/(alpha^T P)^T}
    (:154-156); labels = argmax_k alpha_k P_kb beta_b (:158-160).  alpha [K], beta [B] are carried state.
    Returns (alpha, beta, labels, score) with score [B x K] = (diag(alpha) P diag(beta))^T, the matrix the argmax is
    taken over (so a test can tell real disagreements from near-ties)."""
    x = _f64(logits)
    logp = x - _lse(x, -1)[..., None] if _lse(x, -1).ndim == x.ndim - 1 else x - _lse(x, -1)
    p = np.power(logp, lmbd).T                      # [K, B]
    a = _f64(alpha).reshape(-1, 1).copy()
    b = _f64(beta).reshape(-1, 1).copy()
    for _ in range(int(num_iters)):
        a = 1.0 / (p @ b)
        b = 1.0 / (a.T @ p).T
    score = (a * p * b.T).T                         # [B, K]
    return a[:, 0], b[:, 0], score.argmax(-1), score
