"""Generate golden fixtures by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py      # needs /root/reference (dev container only)

Imports /root/reference/utils/losses.py and the bank classes of models/moco.py,
models/swav.py (faiss is stubbed: it is only pulled in by utils/eval_utils.py:2),
runs forward + torch autograd backward in fp64 where the reference supports it
(BarlowLoss backward is fp32-only: utils/losses.py:139 builds an fp32 eye) on
seeded inputs, and writes inputs + outputs to tests/golden/*.npz.  The GPU box has
no /root/reference: tests only read the committed .npz files.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SSV_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.modules.setdefault("faiss", types.ModuleType("faiss"))
sys.modules.setdefault("wandb", types.ModuleType("wandb"))

from utils import losses as ref_losses  # noqa: E402
from models.moco import MemoryBank as RefMemoryBank  # noqa: E402
from models.swav import FeatureBank as RefFeatureBank, Prototypes as RefPrototypes  # noqa: E402


def randn(seed, *shape, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32).to(dtype)


def clustered(seed, n, d, rho=0.8, dtype=torch.float64):
    """'trained-like' pair: positives correlated rho, a few near-duplicate negatives."""
    a = randn(seed, n, d, dtype=dtype)
    b = rho * a + (1 - rho ** 2) ** 0.5 * randn(seed + 1, n, d, dtype=dtype)
    if n >= 4:
        a[1] = a[0] + 0.01 * randn(seed + 2, d, dtype=dtype)
        b[3] = a[2] + 0.01 * randn(seed + 3, d, dtype=dtype)
    return a, b


def q32(*xs):
    """Round to fp32-representable values so the stored fp32 inputs are exactly what the reference saw."""
    out = tuple(x.float().to(x.dtype) for x in xs)
    return out if len(out) > 1 else out[0]


def npy(t):
    return t.detach().cpu().numpy()


def run(fn, *inputs):
    leaves = [x.clone().requires_grad_(True) for x in inputs]
    loss = fn(*leaves)
    loss.backward()
    return npy(loss), [npy(x.grad) if x.grad is not None else None for x in leaves]


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    torch.manual_seed(0)
    # ---- NT-Xent -------------------------------------------------------------
    cases = {}
    specs = [  # tag, N, d, normalize, tau, clustered
        ("a", 8, 16, True, 0.5, False),
        ("b", 33, 24, False, 1.0, False),
        ("c", 64, 128, True, 0.07, True),
        ("d", 1, 8, True, 0.5, False),
        ("cfg1", 256, 128, True, 0.5, False),   # BASELINE.json configs[0]
    ]
    for tag, n, d, norm, tau, clu in specs:
        if clu:
            zi, zj = clustered(10, n, d)
        else:
            zi, zj = randn(0, n, d), randn(1, n, d)
        if not norm:
            zi, zj = zi * 0.3, zj * 0.3
        zi, zj = q32(zi, zj)
        loss, (gi, gj) = run(ref_losses.SimclrLoss(norm, tau), zi, zj)
        cases.update({f"{tag}_zi": npy(zi).astype(np.float32), f"{tag}_zj": npy(zj).astype(np.float32),
                      f"{tag}_cfg": np.array([float(norm), tau]), f"{tag}_loss": loss,
                      f"{tag}_dzi": gi, f"{tag}_dzj": gj})
    save("ntxent", **cases)

    # ---- MoCo ----------------------------------------------------------------
    cases = {}
    for tag, n, k, d, norm, tau in [("a", 8, 40, 16, True, 1.0), ("b", 32, 1000, 128, True, 0.07),
                                    ("c", 5, 17, 8, False, 0.5)]:
        q, kk = randn(0, n, d), randn(1, n, d)
        mem = torch.nn.functional.normalize(randn(2, k, d), dim=-1)
        mem[: k // 4] = 0.0  # fresh-queue rows are zeros (models/moco.py:26-27)
        if not norm:
            q, kk = q * 0.2, kk * 0.2
        q, kk, mem = q32(q, kk, mem)
        leaves = [q.clone().requires_grad_(True), kk.clone().requires_grad_(True)]
        loss = ref_losses.MocoLoss(norm, tau)(leaves[0], leaves[1], mem)
        loss.backward()
        cases.update({f"{tag}_q": npy(q).astype(np.float32), f"{tag}_k": npy(kk).astype(np.float32),
                      f"{tag}_mem": npy(mem).astype(np.float32), f"{tag}_cfg": np.array([float(norm), tau]),
                      f"{tag}_loss": npy(loss), f"{tag}_dq": npy(leaves[0].grad), f"{tag}_dk": npy(leaves[1].grad)})
    save("moco", **cases)

    # ---- ring buffers ----------------------------------------------------------
    cases = {}
    mb = RefMemoryBank(10, 4)
    fb = RefFeatureBank(7, 3)
    cases["mb_init"] = npy(mb.bank).copy()
    for step, nrows in enumerate([3, 9, 25, 10, 1]):
        batch = randn(100 + step, nrows, 4, dtype=torch.float32)
        if step == 1:
            batch[2] = 0.0  # zero row: normalize(0) == 0
        mb.add_batch(batch)
        cases[f"mb_batch{step}"] = npy(batch)
        cases[f"mb_bank{step}"] = npy(mb.get_vectors()).copy()
        cases[f"mb_ptr{step}"] = np.array(mb.ptr)
        fbatch = randn(200 + step, nrows, 3, dtype=torch.float32)
        fb.add_vectors(fbatch)
        cases[f"fb_batch{step}"] = npy(fbatch)
        cases[f"fb_bank{step}"] = npy(fb.return_vectors("cpu")).copy()
        cases[f"fb_ptr{step}"] = np.array(fb.ptr)
    protos = RefPrototypes(6, 9)
    cases["proto_weight"] = npy(protos.embedding.weight)
    cases["proto_out"] = npy(protos(torch.device("cpu")))
    save("banks", **cases)

    # ---- Barlow (fp32: reference backward is fp32-only) ------------------------
    cases = {}
    for tag, n, d, norm, lm in [("a", 16, 24, True, 0.005), ("b", 32, 40, False, 0.005), ("c", 64, 128, False, 0.02)]:
        g = torch.Generator().manual_seed(7)
        sig = torch.rand(d, generator=g) * 1.5 + 0.5
        mu = torch.randn(d, generator=g)
        zi = randn(0, n, d, dtype=torch.float32) * sig + mu
        zj = (0.7 * zi + 0.5 * randn(1, n, d, dtype=torch.float32) * sig)
        loss, (gi, gj) = run(ref_losses.BarlowLoss(norm, lm), zi, zj)
        cases.update({f"{tag}_zi": npy(zi), f"{tag}_zj": npy(zj), f"{tag}_cfg": np.array([float(norm), lm]),
                      f"{tag}_loss": loss, f"{tag}_dzi": gi, f"{tag}_dzj": gj})
    save("barlow", **cases)

    # ---- SimSiam / BYOL-MSE ------------------------------------------------------
    cases = {}
    o = torch.nn.functional.normalize(randn(0, 12, 20), dim=-1)
    t = torch.nn.functional.normalize(randn(1, 12, 20), dim=-1)
    o, t = q32(o, t)
    loss, (go, gt) = run(ref_losses.SimSiamLoss(), o, t)
    cases.update(dict(o=npy(o).astype(np.float32), t=npy(t).astype(np.float32), ss_loss=loss, ss_do=go, ss_dt=gt))
    loss, (go, gt) = run(torch.nn.MSELoss(), o, t)  # models/byol.py:89
    cases.update(dict(mse_loss=loss, mse_do=go, mse_dt=gt))
    save("rowdot", **cases)

    # ---- ReLIC -------------------------------------------------------------------
    cases = {}
    for tag, n, d, norm, tau, alpha in [("a", 12, 16, True, 1.0, 0.5), ("b", 20, 8, False, 0.7, 0.3),
                                        ("c", 48, 128, True, 0.2, 0.5)]:
        zi, zj, zo = randn(0, n, d), randn(1, n, d), randn(2, n, d)
        if not norm:
            zi, zj, zo = zi * 0.4, zj * 0.4, zo * 0.4
        zi, zj, zo = q32(zi, zj, zo)
        loss, (gi, gj, go_) = run(ref_losses.RelicLoss(norm, tau, alpha), zi, zj, zo)
        cases.update({f"{tag}_zi": npy(zi).astype(np.float32), f"{tag}_zj": npy(zj).astype(np.float32),
                      f"{tag}_zo": npy(zo).astype(np.float32), f"{tag}_cfg": np.array([float(norm), tau, alpha]),
                      f"{tag}_loss": loss, f"{tag}_dzi": gi, f"{tag}_dzj": gj, f"{tag}_dzo": go_})
    save("relic", **cases)

    # ---- Sinkhorn + SwAV -----------------------------------------------------------
    cases = {}
    sw = ref_losses.SwavLoss(0.1, 0.05, 3)
    sw.device = torch.device("cpu")  # utils/losses.py:208 freezes the device at construction
    for tag, b, k, d in [("a", 24, 10, 8), ("b", 64, 30, 16)]:
        z = torch.nn.functional.normalize(randn(0, b, d), dim=-1)
        c = torch.nn.functional.normalize(randn(1, k, d), dim=-1)
        scores = (z @ c.t()).float()
        codes = sw.compute_codes_sinkhorn(scores)
        cases.update({f"sk_{tag}_scores": npy(scores), f"sk_{tag}_codes": npy(codes)})
    for tag, nb, nbank, k, d in [("nobank", 16, 0, 12, 8), ("bank", 16, 20, 12, 8), ("c", 40, 24, 30, 32)]:
        z1 = torch.nn.functional.normalize(randn(0, nb, d, dtype=torch.float32), dim=-1)
        z2 = torch.nn.functional.normalize(0.6 * z1 + 0.4 * randn(1, nb, d, dtype=torch.float32), dim=-1)
        c = torch.nn.functional.normalize(randn(2, k, d, dtype=torch.float32), dim=-1)
        bank = torch.nn.functional.normalize(randn(3, nbank, d, dtype=torch.float32), dim=-1) if nbank else None
        leaves = [z1.clone().requires_grad_(True), z2.clone().requires_grad_(True), c.clone().requires_grad_(True)]
        loss = sw(leaves[0], leaves[1], leaves[2], bank)
        loss.backward()
        cases.update({f"sw_{tag}_z1": npy(z1), f"sw_{tag}_z2": npy(z2), f"sw_{tag}_c": npy(c),
                      f"sw_{tag}_loss": npy(loss), f"sw_{tag}_dz1": npy(leaves[0].grad),
                      f"sw_{tag}_dz2": npy(leaves[1].grad), f"sw_{tag}_dc": npy(leaves[2].grad)})
        if bank is not None:
            cases[f"sw_{tag}_bank"] = npy(bank)
    save("swav", **cases)

    # ---- (f) next rows: DinoLoss, centre EMA, parameter EMA ------------------------------------------
    cases = {}
    dl = ref_losses.DinoLoss()
    for tag, bs, nv, k, ts, tt in [("a", 6, 4, 16, 0.1, 0.04), ("b", 12, 6, 40, 0.1, 0.07)]:
        teacher, student, center = q32(randn(0, bs, 2, k), randn(1, bs, nv, k), 0.1 * randn(2, k))
        st = student.clone().requires_grad_(True)
        loss = dl(teacher, st, ts, tt, center)
        loss.backward()
        cases.update({f"{tag}_teacher": npy(teacher), f"{tag}_student": npy(student), f"{tag}_center": npy(center),
                      f"{tag}_cfg": np.array([ts, tt]), f"{tag}_loss": npy(loss), f"{tag}_dstudent": npy(st.grad)})
    # the reference's own update expressions (models/dino.py:138-141, models/moco.py:110-111), fp32 on CPU
    tf = randn(3, 24, 16, dtype=torch.float32)
    c0 = tf.mean(0)
    tf2 = randn(4, 24, 16, dtype=torch.float32)
    m = 0.9
    c1 = m * c0 + (1 - m) * tf2.mean(0)
    cases.update(center_rows0=npy(tf), center_rows1=npy(tf2), center0=npy(c0), center1=npy(c1), center_m=np.array(m))
    for tag, m in [("p", 0.99), ("q", 0.996)]:
        k_param, q_param = randn(5, 1000, dtype=torch.float32), randn(6, 1000, dtype=torch.float32)
        out = m * k_param + (1.0 - m) * q_param
        cases.update({f"ema_{tag}_t": npy(k_param), f"ema_{tag}_s": npy(q_param), f"ema_{tag}_m": np.array(m),
                      f"ema_{tag}_out": npy(out)})
    # PirlLoss (utils/losses.py:92-117) in fp64 + the per-sample momentum bank of models/pirl.py:22-46 (fp32)
    from models.pirl import MemoryBank as RefPirlBank
    for tag, n, k, d, norm, tau, w in [("pa", 12, 20, 8, True, 0.07, 0.5), ("pb", 24, 50, 16, True, 0.2, 0.3),
                                       ("pc", 10, 16, 8, False, 1.0, 0.5)]:
        img, patch = q32(randn(10, n, d), randn(11, n, d))
        mp = q32(0.7 * torch.nn.functional.normalize(img, dim=-1) + 0.3 * torch.nn.functional.normalize(randn(12, n, d), dim=-1))
        mn = q32(torch.nn.functional.normalize(randn(13, k, d), dim=-1))
        fn = ref_losses.PirlLoss(norm, tau, w)
        li, lp = img.clone().requires_grad_(True), patch.clone().requires_grad_(True)
        loss = fn(li, lp, mp, mn)
        loss.backward()
        cases.update({f"{tag}_img": npy(img), f"{tag}_patch": npy(patch), f"{tag}_mp": npy(mp), f"{tag}_mn": npy(mn),
                      f"{tag}_cfg": np.array([float(norm), tau, w]), f"{tag}_loss": npy(loss), f"{tag}_dimg": npy(li.grad),
                      f"{tag}_dpatch": npy(lp.grad)})
    pb = RefPirlBank(20, 8, momentum=0.5, num_negatives=5)
    idx0 = torch.tensor([3, 7, 0, 19, 11])
    v0, v1 = randn(14, 5, 8, dtype=torch.float32), randn(15, 5, 8, dtype=torch.float32)
    pb.initialize_vectors(idx0, v0)
    cases.update(pbank_idx=idx0.numpy(), pbank_v0=npy(v0), pbank_v1=npy(v1), pbank_after_init=npy(pb.bank.clone()))
    pb.update_vectors(idx0, v1)
    cases.update(pbank_after_update=npy(pb.bank.clone()), pbank_pos=npy(pb.get_positives(torch.tensor([7, 11]))))
    save("next_rows", **cases)

    # ---- SeLA self-labelling: the per-batch body of SeLA.self_label_step (models/sela.py:152-160), the reference's own
    # statements on CPU fp32 tensors (the method itself needs the model and the data loaders).  Case "ref" is the
    # reference's own configuration (lambda = 25, 80 iterations, 128 clusters, batch 500: configs/sela.yaml) - in fp32 it
    # drives alpha to 0 and beta to inf, which is what a drop-in has to reproduce; "s1"/"s2" stay finite.
    import torch.nn.functional as F
    cases = {}
    for tag, b, k, lmbd, iters in [("s1", 50, 16, 3, 20), ("s2", 96, 40, 5, 12), ("ref", 500, 128, 25, 80)]:
        logits = randn(20, b, k, dtype=torch.float32)
        alpha = torch.FloatTensor(k, 1).normal_(0, 1, generator=torch.Generator().manual_seed(21))
        beta = torch.FloatTensor(b, 1).normal_(0, 1, generator=torch.Generator().manual_seed(22))
        a, bb = alpha.clone(), beta.clone()
        log_probs = torch.pow(F.log_softmax(logits, -1), lmbd).t()                         # sela.py:152
        for _ in range(iters):
            a = 1.0 / torch.mm(log_probs, bb)                                               # sela.py:155
            bb = 1.0 / torch.mm(a.t(), log_probs).t()                                       # sela.py:156
        alpha_diag = torch.eye(a.size(0)) * a                                               # sela.py:158
        beta_diag = torch.eye(bb.size(0)) * bb                                              # sela.py:159
        score = (alpha_diag @ log_probs @ beta_diag).t()
        labels = score.argmax(-1)                                                           # sela.py:160
        cases.update({f"{tag}_logits": npy(logits), f"{tag}_alpha0": npy(alpha), f"{tag}_beta0": npy(beta),
                      f"{tag}_cfg": np.array([lmbd, iters]), f"{tag}_alpha": npy(a), f"{tag}_beta": npy(bb),
                      f"{tag}_labels": labels.numpy()})
        if tag != "ref":   # (all NaN there)
            cases[f"{tag}_score"] = npy(score)
    save("sela", **cases)


if __name__ == "__main__":
    main()
