"""Pin the CPU oracle (oracle/ssl_oracle.py) against outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2, rel_scalar
from oracle import ssl_oracle as O

TOL64 = 1e-9   # fixtures computed by the reference in fp64
TOL32 = 2e-5   # fixtures computed by the reference in fp32 (Barlow, SwAV, Sinkhorn)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d", "cfg1"])
def test_ntxent(tag):
    g = load_golden("ntxent")
    norm, tau = bool(g[f"{tag}_cfg"][0]), float(g[f"{tag}_cfg"][1])
    loss, dzi, dzj = O.ntxent(g[f"{tag}_zi"], g[f"{tag}_zj"], norm, tau)
    if tag == "d":  # N=1: a single negative-free row pair, loss = 0, grads = 0
        assert abs(loss - float(g[f"{tag}_loss"])) < 1e-12
        assert np.abs(dzi - g[f"{tag}_dzi"]).max() < 1e-12
        return
    assert rel_scalar(loss, g[f"{tag}_loss"]) < TOL64
    assert rel_l2(dzi, g[f"{tag}_dzi"]) < TOL64
    assert rel_l2(dzj, g[f"{tag}_dzj"]) < TOL64


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_moco(tag):
    g = load_golden("moco")
    norm, tau = bool(g[f"{tag}_cfg"][0]), float(g[f"{tag}_cfg"][1])
    loss, dq, dk = O.moco(g[f"{tag}_q"], g[f"{tag}_k"], g[f"{tag}_mem"], norm, tau)
    assert rel_scalar(loss, g[f"{tag}_loss"]) < TOL64
    assert rel_l2(dq, g[f"{tag}_dq"]) < TOL64
    assert rel_l2(dk, g[f"{tag}_dk"]) < TOL64


def test_ring_buffers_bit_exact():
    g = load_golden("banks")
    bank, ptr = g["mb_init"].copy(), 0
    assert not bank.any()
    fbank, fptr = np.zeros((7, 3), np.float32), 0
    for step in range(5):
        bank, ptr = O.ring_enqueue(bank, ptr, g[f"mb_batch{step}"], normalize=True)
        assert ptr == int(g[f"mb_ptr{step}"])
        # which rows were overwritten is exact; values to 1 ulp of fp32 (reduction order)
        np.testing.assert_allclose(bank, g[f"mb_bank{step}"], rtol=2e-7, atol=0)
        assert ((bank == 0) == (g[f"mb_bank{step}"] == 0)).all()
        fbank, fptr = O.ring_enqueue(fbank, fptr, g[f"fb_batch{step}"], normalize=False)
        assert fptr == int(g[f"fb_ptr{step}"])
        assert np.array_equal(fbank, g[f"fb_bank{step}"])
    assert rel_l2(O.prototypes_forward(g["proto_weight"]), g["proto_out"]) < 1e-6


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_barlow(tag):
    g = load_golden("barlow")
    norm, lm = bool(g[f"{tag}_cfg"][0]), float(g[f"{tag}_cfg"][1])
    loss, dzi, dzj = O.barlow(g[f"{tag}_zi"], g[f"{tag}_zj"], norm, lm)
    assert rel_scalar(loss, g[f"{tag}_loss"]) < TOL32
    assert rel_l2(dzi, g[f"{tag}_dzi"]) < TOL32
    assert rel_l2(dzj, g[f"{tag}_dzj"]) < TOL32


def test_rowdot():
    g = load_golden("rowdot")
    loss, do, dt = O.simsiam(g["o"], g["t"])
    assert rel_scalar(loss, g["ss_loss"]) < TOL64
    assert rel_l2(do, g["ss_do"]) < TOL64 and rel_l2(dt, g["ss_dt"]) < TOL64
    loss, do, dt = O.mse(g["o"], g["t"])
    assert rel_scalar(loss, g["mse_loss"]) < TOL64
    assert rel_l2(do, g["mse_do"]) < TOL64 and rel_l2(dt, g["mse_dt"]) < TOL64


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_relic(tag):
    g = load_golden("relic")
    norm, tau, alpha = bool(g[f"{tag}_cfg"][0]), float(g[f"{tag}_cfg"][1]), float(g[f"{tag}_cfg"][2])
    loss, dzi, dzj, dzo = O.relic(g[f"{tag}_zi"], g[f"{tag}_zj"], g[f"{tag}_zo"], norm, tau, alpha)
    assert rel_scalar(loss, g[f"{tag}_loss"]) < TOL64
    assert rel_l2(dzi, g[f"{tag}_dzi"]) < TOL64
    assert rel_l2(dzj, g[f"{tag}_dzj"]) < TOL64
    assert rel_l2(dzo, g[f"{tag}_dzo"]) < TOL64


@pytest.mark.parametrize("tag", ["a", "b"])
def test_sinkhorn(tag):
    g = load_golden("swav")
    codes = O.sinkhorn(g[f"sk_{tag}_scores"], 0.05, 3)
    assert rel_l2(codes, g[f"sk_{tag}_codes"]) < TOL32
    np.testing.assert_allclose(codes.sum(1), 1.0, rtol=1e-12)


@pytest.mark.parametrize("tag", ["nobank", "bank", "c"])
def test_swav(tag):
    g = load_golden("swav")
    bank = g.get(f"sw_{tag}_bank")
    loss, dz1, dz2, dc = O.swav(g[f"sw_{tag}_z1"], g[f"sw_{tag}_z2"], g[f"sw_{tag}_c"], bank, 0.1, 0.05, 3)
    assert rel_scalar(loss, g[f"sw_{tag}_loss"]) < TOL32
    assert rel_l2(dz1, g[f"sw_{tag}_dz1"]) < 5 * TOL32
    assert rel_l2(dz2, g[f"sw_{tag}_dz2"]) < 5 * TOL32
    assert rel_l2(dc, g[f"sw_{tag}_dc"]) < 5 * TOL32


# --------------------------------------------------------------------------- (f) next rows
@pytest.mark.parametrize("tag", ["a", "b"])
def test_dino_oracle_vs_reference(tag):
    g = load_golden("next_rows")
    ts, tt = (float(x) for x in g[f"{tag}_cfg"])
    loss, ds = O.dino(g[f"{tag}_teacher"], g[f"{tag}_student"], ts, tt, g[f"{tag}_center"])
    assert rel_scalar(loss, float(g[f"{tag}_loss"])) < 1e-9
    assert rel_l2(ds, g[f"{tag}_dstudent"]) < 1e-9


def test_ema_and_center_oracle_bit_exact_vs_reference():
    g = load_golden("next_rows")
    for tag in ("p", "q"):
        out = O.ema_update(g[f"ema_{tag}_t"], g[f"ema_{tag}_s"], float(g[f"ema_{tag}_m"]))
        assert out.dtype == np.float32 and np.array_equal(out, g[f"ema_{tag}_out"])
    c0 = O.dino_center_update(None, g["center_rows0"], float(g["center_m"]))
    np.testing.assert_allclose(c0, g["center0"], rtol=2e-6, atol=1e-7)  # torch / numpy sum the 24 rows in different orders
    c1 = O.dino_center_update(g["center0"], g["center_rows1"], float(g["center_m"]))
    np.testing.assert_allclose(c1, g["center1"], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("tag", ["pa", "pb", "pc"])
def test_pirl_oracle_vs_reference(tag):
    g = load_golden("next_rows")
    norm, tau, w = bool(g[f"{tag}_cfg"][0]), float(g[f"{tag}_cfg"][1]), float(g[f"{tag}_cfg"][2])
    loss, dimg, dpatch = O.pirl(g[f"{tag}_img"], g[f"{tag}_patch"], g[f"{tag}_mp"], g[f"{tag}_mn"], norm, tau, w)
    assert rel_scalar(loss, float(g[f"{tag}_loss"])) < TOL64
    assert rel_l2(dimg, g[f"{tag}_dimg"]) < TOL64
    assert rel_l2(dpatch, g[f"{tag}_dpatch"]) < TOL64


def test_pirl_bank_oracle_vs_reference():
    g = load_golden("next_rows")
    bank = O.pirl_bank_update(np.zeros((20, 8), np.float32), g["pbank_idx"], g["pbank_v0"])
    np.testing.assert_allclose(bank, g["pbank_after_init"], rtol=3e-7, atol=0)
    bank2 = O.pirl_bank_update(g["pbank_after_init"], g["pbank_idx"], g["pbank_v1"], 0.5)
    np.testing.assert_allclose(bank2, g["pbank_after_update"], rtol=3e-7, atol=0)
    assert np.array_equal(g["pbank_pos"], g["pbank_after_update"][[7, 11]])


@pytest.mark.parametrize("tag", ["s1", "s2"])
def test_sela_self_label_oracle_vs_reference(tag):
    """SeLA self-labelling (models/sela.py:152-160): the fp64 oracle against the reference's own fp32 statements.
    alpha / beta individually carry a gauge (alpha * c, beta / c) that the fp32 iteration fixes by its rounding history,
    so the comparison is on the gauge-free score matrix and on the labels (wherever the top-2 margin is not a near-tie)."""
    g = load_golden("sela")
    lmbd, iters = (int(v) for v in g[f"{tag}_cfg"])
    a, b, labels, score = O.sela_self_label(g[f"{tag}_logits"], g[f"{tag}_alpha0"][:, 0], g[f"{tag}_beta0"][:, 0], lmbd, iters)
    ref_score = g[f"{tag}_score"].astype(np.float64)
    np.testing.assert_allclose(score, ref_score, rtol=5e-3, atol=1e-6 * np.abs(ref_score).max())
    top2 = np.sort(ref_score, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-3 * np.abs(top2[:, 1])
    assert clear.mean() > 0.9
    assert np.array_equal(labels[clear], g[f"{tag}_labels"][clear])


def test_sela_reference_configuration_degenerates_in_fp32():
    """The reference's own configuration (lambda 25, 80 iterations, 128 clusters: configs/sela.yaml) underflows in fp32:
    alpha -> 0, beta -> inf, every label 0.  The fixture pins that behaviour (a drop-in has to reproduce it)."""
    g = load_golden("sela")
    assert not np.any(g["ref_alpha"]) and np.all(np.isinf(g["ref_beta"])) and not np.any(g["ref_labels"])
